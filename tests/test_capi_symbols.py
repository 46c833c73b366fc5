"""The C-ABI library loads, exports every symbol include/kmc_b200.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kmc_b200.h")).read()
    return sorted(set(re.findall(r"KMC_B200_API\s+[\w\s\*]+?\b(kmc_b200_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("kmc_b200_deskew_frame_device", "kmc_b200_deskew_batch_device", "kmc_b200_deskew_frame_host",
                 "kmc_b200_deskew_batch_host", "kmc_b200_deskew_batch_multi_gpu", "kmc_b200_frame_params_from_poses",
                 "kmc_b200_last_error", "kmc_b200_pseudo_time_stamps_device", "kmc_b200_handle_create"):
        assert must in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol(capi):
    lib = capi.lib()
    exported = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (kmc_b200_\w+)", exported))
    declared = declared_symbols()
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert exported <= set(declared), f"exported but undeclared: {sorted(exported - set(declared))}"
    for name in declared:
        assert hasattr(lib, name)
    assert set(capi.SIGNATURES) == set(declared)


def test_header_compiles_as_plain_c():
    """No C++/torch/Eigen types leak into the boundary: the header must be valid C99."""
    src = '#include "kmc_b200.h"\nint main(void) { kmc_b200_frame_params p; return (int)sizeof(p) - 64; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-x", "c", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_version_and_status_strings(capi):
    lib = capi.lib()
    assert lib.kmc_b200_version() == 200
    assert lib.kmc_b200_status_string(0) == b"ok"
    assert b"interval" in lib.kmc_b200_status_string(capi.ERR_TIME_OUT_OF_RANGE)
    assert C.sizeof(capi.FrameParams) == 64


def test_argument_errors_are_statuses(capi):
    lib = capi.lib()
    p = capi.frame_params_from_twist([1, 0, 0, 0, 0, 0.01], 0.5)
    assert lib.kmc_b200_deskew_frame_device(None, None, -1, C.byref(p), 0, None) == capi.ERR_BAD_SIZE
    assert lib.kmc_b200_deskew_frame_device(None, None, 10, C.byref(p), 7, None) == capi.ERR_BAD_MODE
    assert lib.kmc_b200_deskew_frame_device(None, None, 10, C.byref(p), 0, None) == capi.ERR_NULL_POINTER
    assert lib.kmc_b200_deskew_frame_device(None, None, 10, None, 0, None) == capi.ERR_NULL_POINTER
    assert lib.kmc_b200_deskew_frame_device(16, 24, 10, C.byref(p), 0, None) == capi.ERR_BAD_SIZE  # misaligned
    assert lib.kmc_b200_deskew_frame_device(None, None, 0, C.byref(p), 0, None) == capi.OK  # empty scan
    assert lib.kmc_b200_deskew_batch_device(None, None, None, None, 0, 0, 0, None) == capi.OK  # empty batch
    assert lib.kmc_b200_deskew_batch_device(None, None, None, None, 3, 10, 0, None) == capi.ERR_NULL_POINTER
    assert lib.kmc_b200_frame_params_from_twist(None, 0.5, C.byref(p)) == capi.ERR_NULL_POINTER
    assert lib.kmc_b200_deskew_frame_host(None, None, None, 5, C.byref(p), 0) == capi.ERR_NULL_POINTER
    assert capi.last_error() != ""


def test_no_cpu_fallback_without_a_gpu(capi):
    """On a box without a CUDA device every compute entry point must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present; the no-device behaviour is exercised on the CPU box")
    lib = capi.lib()
    assert lib.kmc_b200_device_count() < 0
    pts = np.zeros((8, 4), dtype=np.float32)
    out = np.full((8, 4), 7.0, dtype=np.float32)
    p = capi.frame_params_from_twist([1, 0, 0, 0, 0, 0.01], 0.5)
    rc = lib.kmc_b200_deskew_frame_device(pts.ctypes.data, out.ctypes.data, 8, C.byref(p), 0, None)
    assert rc in (capi.ERR_CUDA, capi.ERR_NO_DEVICE)
    assert np.all(out == 7.0)
    with pytest.raises(capi.KmcError) as e:
        capi.Handle(0, 1024)
    assert e.value.status in (capi.ERR_CUDA, capi.ERR_NO_DEVICE)


def test_missing_library_raises(monkeypatch, capi):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libkmc_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        capi.lib()


def test_bench_reference_arm_is_hermetic_and_draws_the_same_twists():
    """bench.py --impl reference generates its scans and twists with numpy and never imports the CUDA library; its twist
    generator is the restatement of kmc_b200_synth_frame_params."""
    import importlib.util
    import os
    import subprocess
    import sys
    import numpy as np
    from kitti_motion_compensation_b200 import capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for first in (0, 12_345):
        _, xi = capi.synth_frame_params(50, bench.SEED, first, 0.5)
        assert np.array_equal(bench.numpy_twists(50, bench.SEED, first), xi)
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--points', '2000'];\n"
            "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libkmc_b200' not in maps and 'kitti_motion_compensation_lib' not in maps, 'reference arm mapped the CUDA library'\n"
            "assert 'kitti_motion_compensation_b200' not in ' '.join(sys.modules), 'reference arm imported the package'\n"
            "print('hermetic')")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "hermetic" in r.stdout, r.stderr[-2000:]
    assert '"impl": "reference"' in r.stdout
