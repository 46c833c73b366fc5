"""Builds libkmc_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

The library is self-contained (static cudart) so that it loads in a plain `ctypes.CDLL`, travels to the GPU box with
the repository snapshot and shows up as an in-tree native library in the loaded-object list.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libkmc_b200.so")
DROPIN_LIB_PATH = os.path.join(LIB_DIR, "libkitti_motion_compensation_lib.so")

SOURCES = ["kmc_kernels.cu", "kmc_kernels_bulk.cu", "kmc_capi.cu", "kmc_pipeline.cu", "kmc_host_math.cpp", "kmc_run.cpp"]
HEADERS = ["kmc_kernels.cuh", "kmc_point_math.cuh", "kmc_host_math.hpp", "kmc_internal.hpp", "kmc_host_pool.hpp", os.path.join(REPO_DIR, "include", "kmc_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build kitti_motion_compensation_b200")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not _stale(LIB_PATH, deps):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    # every source is an independent translation unit (no -rdc): compile them side by side, then link
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f not in ("--shared", "-cudart", "static")]
    includes = ["-I", os.path.join(REPO_DIR, "include"), "-I", CSRC]

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, os.path.basename(src) + ".o")
        cmd = [_nvcc(), *compile_flags, *includes, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:
        objs = list(pool.map(compile_one, srcs))
    subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static", "-o", LIB_PATH, *objs], check=True)
    shutil.rmtree(obj_dir, ignore_errors=True)  # every build recompiles all sources; the objects need not travel to the GPU box
    return LIB_PATH


DROPIN_HEADERS = ["data_types.hpp", "eigen_shim.hpp", "lie_algebra.hpp", "trajectory_interpolation.hpp", "timestamp_mocking.hpp",
                  "motion_compensation.hpp", "camera_model.hpp", "data_io.hpp", "handlers.hpp", "utils.hpp", "data_handle.hpp"]


def _cxx() -> str:
    return os.environ.get("CXX") or shutil.which("g++") or "g++"


def build_dropin(force: bool = False) -> str:
    """libkitti_motion_compensation_lib.so — the C++ mirror of the reference API (reference CMakeLists.txt:27-29 names
    its library the same), linked against libkmc_b200.so next to it."""
    build()  # never forced from here: a forced caller has just rebuilt libkmc_b200.so itself
    src = os.path.join(CSRC, "kmc_dropin.cpp")
    inc = os.path.join(REPO_DIR, "include")
    deps = [src, os.path.join(inc, "kmc_b200.h"), os.path.abspath(__file__)]
    deps += [os.path.join(inc, "kitti_motion_compensation", h) for h in DROPIN_HEADERS]
    if not force and not _stale(DROPIN_LIB_PATH, deps):
        return DROPIN_LIB_PATH
    cmd = [_cxx(), "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-I", inc, "-o", DROPIN_LIB_PATH, src,
           "-L", LIB_DIR, "-lkmc_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return DROPIN_LIB_PATH


def build_cpp_test(name: str, force: bool = False) -> str:
    """Compiles tests/cpp/<name>.cpp against the drop-in library; returns the binary path (tests/cpp/_build/<name>)."""
    build_dropin()
    tests = os.path.join(REPO_DIR, "tests", "cpp")
    out_dir = os.path.join(tests, "_build")
    os.makedirs(out_dir, exist_ok=True)
    src, out = os.path.join(tests, name + ".cpp"), os.path.join(out_dir, name)
    if not force and not _stale(out, [src, os.path.join(tests, "mini_gtest.hpp"), DROPIN_LIB_PATH]):
        return out
    cmd = [_cxx(), "-std=c++17", "-O1", "-Wall", "-I", os.path.join(REPO_DIR, "include"), "-I", tests, "-o", out, src,
           "-L", LIB_DIR, "-lkitti_motion_compensation_lib", "-lkmc_b200", "-lpthread", f"-Wl,-rpath,{LIB_DIR}"]
    subprocess.run(cmd, check=True)
    return out


EXAMPLES = ["motion_compensate_runs", "bench_motion_compensate_frame"]


def build_example(force: bool = False, name: str | None = None) -> str:
    """examples/<name>.cpp -> kitti_motion_compensation_b200/lib/<name> (all of EXAMPLES when name is None; returns the
    path of motion_compensate_runs, the reference's CLI, in that case)."""
    build_dropin()
    if name is None:
        paths = [build_example(force, n) for n in EXAMPLES]
        return paths[0]
    src = os.path.join(REPO_DIR, "examples", name + ".cpp")
    out = os.path.join(LIB_DIR, name)
    if not force and not _stale(out, [src, DROPIN_LIB_PATH]):
        return out
    cmd = [_cxx(), "-std=c++17", "-O2", "-Wall", "-I", os.path.join(REPO_DIR, "include"), "-o", out, src,
           "-L", LIB_DIR, "-lkitti_motion_compensation_lib", "-lkmc_b200", "-lpthread", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return out


REFERENCE_DIR = os.environ.get("KMC_REFERENCE_DIR", "/root/reference")
REFERENCE_TESTS = ["test_motion_compensation", "test_timestamp_mocking", "test_lie_algebra", "test_trajectory_interpolation",
                   "test_oxts_to_pose"]
REF_TEST_ROOT = os.path.join(REPO_DIR, "tests", "cpp", "_ref_build")


def build_reference_tests(force: bool = False) -> list[str]:
    """Compiles the reference's OWN gtest files — unmodified, where they lie under /root/reference/test — against this
    repository's include/ (the drop-in headers) and libkitti_motion_compensation_lib.so, with <gtest/gtest.h> supplied by
    tests/cpp/gtest_stub (GoogleTest is not in the image).  Binaries go to tests/cpp/_ref_build/build/ and the small run
    folder the tests open as "../testing_assets/..." is copied beside them (images left out), so that the binaries run on
    the GPU box, where /root/reference does not exist.  Nothing here is tracked by git.  Returns the binaries that exist."""
    build_dropin()
    out_dir = os.path.join(REF_TEST_ROOT, "build")
    src_dir = os.path.join(REFERENCE_DIR, "test")
    if os.path.isdir(src_dir):
        os.makedirs(out_dir, exist_ok=True)
        assets_src = os.path.join(REFERENCE_DIR, "testing_assets")
        assets_dst = os.path.join(REF_TEST_ROOT, "testing_assets")
        if os.path.isdir(assets_src) and (force or not os.path.isdir(assets_dst)):
            shutil.rmtree(assets_dst, ignore_errors=True)
            shutil.copytree(assets_src, assets_dst, ignore=shutil.ignore_patterns("image_0*"))
        tests = os.path.join(REPO_DIR, "tests", "cpp")
        for name in REFERENCE_TESTS:
            src, out = os.path.join(src_dir, name + ".cpp"), os.path.join(out_dir, name)
            if not os.path.exists(src):
                continue
            if not force and not _stale(out, [src, os.path.join(tests, "mini_gtest.hpp"), DROPIN_LIB_PATH]):
                continue
            cmd = [_cxx(), "-std=c++17", "-O1", "-I", os.path.join(tests, "gtest_stub"), "-I", os.path.join(REPO_DIR, "include"),
                   "-o", out, src, "-L", LIB_DIR, "-lkitti_motion_compensation_lib", "-lkmc_b200", "-lpthread",
                   "-Wl,-rpath,$ORIGIN/../../../../kitti_motion_compensation_b200/lib"]
            subprocess.run(cmd, check=True)
    return [os.path.join(out_dir, n) for n in REFERENCE_TESTS if os.path.exists(os.path.join(out_dir, n))]


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_dropin(force="--force" in sys.argv))
    print(build_example(force="--force" in sys.argv))
    print(build_reference_tests(force="--force" in sys.argv))
