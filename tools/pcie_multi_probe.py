#!/usr/bin/env python
"""Why does the end-to-end (host buffer) throughput not scale with the number of GPUs?  One process per GPU, all copying at
the same time: H2D only, D2H only, both directions — unbound, and with each process (and therefore its pinned buffers,
first-touched after binding) bound to the CPUs of the NUMA node its GPU hangs off (/sys/bus/pci/devices/<bdf>/numa_node).
Prints per-GPU and aggregate GB/s per direction.  Context for DESIGN.md §7; not a bench.py number.

  python tools/pcie_multi_probe.py --gpus 2
"""
import argparse
import glob
import os
import re
import subprocess
import time

import torch
import torch.multiprocessing as mp


def node_cpus(node):
    out = []
    try:
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            out.extend(range(int(a), int(b or a) + 1))
    except OSError:
        pass
    return out


def gpu_numa_node(index):
    q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)], capture_output=True, text=True)
    bdf = q.stdout.strip().lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    try:
        return int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read()), bdf
    except (OSError, ValueError):
        return -1, bdf


def worker(rank, world, bind, seconds, n_bytes, barrier, results):
    torch.cuda.set_device(rank)
    node, bdf = gpu_numa_node(rank)
    note = "unbound"
    if bind == "local" and node >= 0 and node_cpus(node):
        os.sched_setaffinity(0, node_cpus(node))
        note = f"node{node}"
    elif bind == "remote":
        nodes = sorted(int(re.findall(r"\d+$", p)[0]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
        other = [n for n in nodes if n != node]
        if other and node_cpus(other[0]):
            os.sched_setaffinity(0, node_cpus(other[0]))
            note = f"node{other[0]} (remote)"
    dev_a = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    dev_b = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    pin_a = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    pin_b = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    pin_a.fill_(1)
    pin_b.fill_(2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            dev_a.copy_(pin_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            pin_b.copy_(dev_b, non_blocking=True)

    out = {"gpu": rank, "bdf": bdf, "gpu_node": node, "bound": note}
    for name, fns in (("h2d", (h2d,)), ("d2h", (d2h,)), ("both", (h2d, d2h))):
        for f in fns:
            f()
        torch.cuda.synchronize()
        barrier.wait(timeout=60)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < seconds:
            for f in fns:
                f()
            torch.cuda.synchronize()
            reps += 1
        out[name] = reps * n_bytes / (time.perf_counter() - t0) / 1e9
        barrier.wait(timeout=60)
    results.put(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--mbytes", type=int, default=512)
    args = ap.parse_args()
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'; free -g | head -2", shell=True, capture_output=True, text=True).stdout)
    mp.set_start_method("spawn", force=True)
    for world in sorted({1, args.gpus}):
        n_nodes = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
        for bind in (("none", "local", "remote") if n_nodes > 1 else ("none",)):
            barrier = mp.Barrier(world)
            results = mp.Queue()
            procs = [mp.Process(target=worker, args=(r, world, bind, args.seconds, args.mbytes << 20, barrier, results)) for r in range(world)]
            for p in procs:
                p.start()
            try:
                rows = sorted((results.get(timeout=60 + 8 * args.seconds) for _ in procs), key=lambda r: r["gpu"])
            except Exception as e:  # a worker died: do not wait for it
                for p in procs:
                    p.terminate()
                print(f"{world} GPU(s), bind={bind}: worker failed ({type(e).__name__})", flush=True)
                continue
            for p in procs:
                p.join(timeout=30)
            agg = {k: sum(r[k] for r in rows) for k in ("h2d", "d2h", "both")}
            print(f"{world} GPU(s), bind={bind:6s}: aggregate H2D {agg['h2d']:6.1f}  D2H {agg['d2h']:6.1f}  both {agg['both']:6.1f} GB/s each way   per GPU: "
                  + "  ".join(f"[{r['gpu']} {r['bdf']} gpu_node={r['gpu_node']} {r['bound']}: {r['h2d']:.1f}/{r['d2h']:.1f}/{r['both']:.1f}]" for r in rows), flush=True)


if __name__ == "__main__":
    main()
