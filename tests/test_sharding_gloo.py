"""Multi-GPU host logic on CPU: world_size-2 gloo run of the same sharding + max-over-ranks reduction bench.py uses.
Frames are independent, so the N>1 path has NO data-path collective; what needs testing is that the shards tile the
batch exactly, that per-frame constants do not depend on the shard they were generated in, and that the throughput
reduction (sum of points, max of time) is what bench.py prints."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kitti_motion_compensation_b200 import capi
    b, e = capi.shard_range(n_frames, world, rank)
    params, xi = capi.synth_frame_params(e - b, 20110926, b, 0.5)  # strong scaling: shard of one batch
    np.save(os.path.join(out_dir, f"params_{rank}.npy"), params.view(np.uint8))
    # what bench.py reduces: total points (SUM) and elapsed time (MAX)
    pts = torch.tensor([float((e - b) * 130_000)], dtype=torch.float64)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(pts, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ranges = [None] * world
    dist.all_gather_object(ranges, (b, e))
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), np.array([pts.item(), t.item()] + [v for r in ranges for v in r]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction(tmp_path):
    sys.path.insert(0, ROOT)
    from kitti_motion_compensation_b200 import build, capi
    build.build()
    world, n_frames = 2, 10_001
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    reduced = np.load(tmp_path / "reduced.npy")
    assert reduced[0] == n_frames * 130_000 and reduced[1] == 2.0
    ranges = reduced[2:].astype(int).reshape(world, 2)
    assert ranges[0, 0] == 0 and ranges[-1, 1] == n_frames and ranges[0, 1] == ranges[1, 0]
    # concatenated shard tables == the single-rank table: a frame does not know which GPU it runs on
    whole, _ = capi.synth_frame_params(n_frames, 20110926, 0, 0.5)
    parts = np.concatenate([np.load(tmp_path / f"params_{r}.npy") for r in range(world)])
    assert parts.tobytes() == whole.view(np.uint8).tobytes()


def _checksum_worker(rank, world, port, n_frames, corrupt_rank, out_dir):
    """What bench.py's strong-scaling leg does after its timed region, with numpy standing in for the GPU: every rank
    checksums the frames of ITS shard, the checksums are all-gathered (padded to a common length), rank 0 compares them with
    the checksums of the whole batch computed in one piece."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib.util
    from kitti_motion_compensation_b200 import capi
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    points = 257

    def frame(f):  # the "deskewed output" of frame f: depends on the frame only, never on the rank that produced it
        return np.random.default_rng(1000 + f).standard_normal((points, 4)).astype(np.float32)

    b, e = capi.shard_range(n_frames, world, rank)
    shard = np.concatenate([frame(f) for f in range(b, e)]) if e > b else np.zeros((0, 4), np.float32)
    if rank == corrupt_rank and e > b:
        shard.view(np.uint32)[points * ((e - b) // 2) + 5, 1] ^= 1  # one flipped bit in one frame of this rank's shard
    sums = capi.frame_checksums_numpy(shard, np.arange(0, (e - b + 1) * points, points))
    cap = -(-n_frames // world)
    mine = torch.zeros(cap, dtype=torch.int64)
    mine[: e - b] = torch.from_numpy(sums.view(np.int64))
    gathered = [torch.zeros(cap, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        whole = np.concatenate([frame(f) for f in range(n_frames)])
        whole_sums = torch.from_numpy(capi.frame_checksums_numpy(whole, np.arange(0, (n_frames + 1) * points, points)).view(np.int64))
        equal, compared = bench.shards_bit_equal(gathered, whole_sums, n_frames, world, capi.shard_range)
        np.save(os.path.join(out_dir, f"equal_{corrupt_rank}.npy"), np.array([int(equal), compared]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_checksums_compare_like_the_strong_leg(tmp_path):
    """bench.py's `shards_bit_equal` over a gloo all-gather: equal when every rank holds the right bits, false when one bit of
    one frame on one rank differs (frame counts that do not divide evenly included)."""
    sys.path.insert(0, ROOT)
    from kitti_motion_compensation_b200 import build
    build.build()
    world, n_frames = 2, 11
    for corrupt_rank in (-1, 1):
        mp.spawn(_checksum_worker, args=(world, _free_port(), n_frames, corrupt_rank, str(tmp_path)), nprocs=world, join=True)
        equal, compared = np.load(tmp_path / f"equal_{corrupt_rank}.npy")
        assert compared == n_frames
        assert bool(equal) == (corrupt_rank < 0)


def test_bench_reference_arm_runs_on_cpu_and_prints_one_json_line():
    """bench.py --impl reference (the oracle port on the host cores) must work without a GPU and under torchrun env."""
    import json
    import subprocess
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--points", "20000"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "Mpoints/s" and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] in ("reference", "port") and rec["cpu_baseline"]["cores"] >= 1  # "reference" when oracle/_ref is built
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["d2h_bytes_per_step"] == 0
    # the other ranks exit 0 without work
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
