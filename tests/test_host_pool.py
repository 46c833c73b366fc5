"""The library's host thread pool (csrc/kmc_host_pool.hpp) as a plain C++ unit: compiled with g++ here, with ThreadSanitizer when
the toolchain links it.  The pool carries the host passes of kmc::MotionCompensateFrame and the staging of pageable buffers."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_host_pool.cpp")
INC = os.path.join(ROOT, "kitti_motion_compensation_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "cpp", "_build")


def build(flags, name):
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, name)
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", *flags, "-I", INC, "-o", exe, SRC, "-lpthread"], capture_output=True, text=True)
    return exe if r.returncode == 0 else None, r.stderr


def test_host_pool_runs_every_block_exactly_once():
    exe, err = build([], "test_host_pool")
    assert exe, err
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "host pool ok" in r.stdout, r.stdout + r.stderr


def test_host_pool_under_thread_sanitizer():
    import pytest
    exe, err = build(["-fsanitize=thread", "-g"], "test_host_pool_tsan")
    if not exe:
        pytest.skip("this toolchain cannot link -fsanitize=thread: " + err.strip().splitlines()[-1][:200])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env={**os.environ, "TSAN_OPTIONS": "halt_on_error=1"})
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container (memory mapping)")
    assert r.returncode == 0 and "host pool ok" in r.stdout and "WARNING: ThreadSanitizer" not in r.stderr, r.stdout + r.stderr[-3000:]
