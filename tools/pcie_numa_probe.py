#!/usr/bin/env python
"""Does the placement of pinned host memory (NUMA node of the allocating thread) change the PCIe ceilings of GPU 0?

For every NUMA node the box exposes: bind this process to the node's CPUs, allocate + first-touch pinned buffers, measure
H2D alone, D2H alone and both at once (plain cudaMemcpyAsync through torch, 1 GiB each way).  Prints one line per node.
Context for DESIGN.md §4 (the end-to-end number is PCIe-bound); not a bench.py number.
"""
import glob
import os
import re
import subprocess
import time

import torch


def cpus_of(node_dir):
    out = []
    for part in open(os.path.join(node_dir, "cpulist")).read().strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def measure(n_bytes=1 << 30, reps=5):
    torch.cuda.set_device(0)
    dev_a = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    dev_b = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    pin_a = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    pin_b = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    pin_a.fill_(1)
    pin_b.fill_(2)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return reps * n_bytes / (time.perf_counter() - t0) / 1e9

    def both():
        with torch.cuda.stream(s1):
            dev_a.copy_(pin_a, non_blocking=True)
        with torch.cuda.stream(s2):
            pin_b.copy_(dev_b, non_blocking=True)

    h2d = timed(lambda: dev_a.copy_(pin_a, non_blocking=True))
    d2h = timed(lambda: pin_b.copy_(dev_b, non_blocking=True))
    bi = timed(both)
    return h2d, d2h, bi


def main():
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=60).stdout)
    except Exception as e:  # noqa: BLE001
        print("nvidia-smi topo failed:", e)
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"), key=lambda p: int(re.findall(r"\d+$", p)[0]))
    print(f"{len(nodes)} NUMA node(s); {os.cpu_count()} CPUs; affinity {sorted(os.sched_getaffinity(0))}", flush=True)
    all_cpus = sorted(os.sched_getaffinity(0))
    h2d, d2h, bi = measure()
    print(f"unbound            : H2D {h2d:5.1f}  D2H {d2h:5.1f}  both {bi:5.1f} GB/s each way", flush=True)
    for nd in nodes:
        cpus = [c for c in cpus_of(nd) if c in all_cpus]
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        h2d, d2h, bi = measure()
        print(f"{os.path.basename(nd):8s} cpus {cpus[0]}-{cpus[-1]}: H2D {h2d:5.1f}  D2H {d2h:5.1f}  both {bi:5.1f} GB/s each way", flush=True)
    os.sched_setaffinity(0, all_cpus)


if __name__ == "__main__":
    main()
