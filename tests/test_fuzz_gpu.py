"""Property-based stress of the batch paths (register and TMA-staged kernels, device and host entry points): random
frame sizes (empties, singletons, tile-sized, odd), random twists and requested fractions, random staging capacities.
Every point of every frame must match the double-precision closed form, the w lane must pass through, and nothing may
be written outside the batch.  `-m gpu`."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import helpers
from helpers import TOL_M

pytestmark = pytest.mark.gpu

SIZES = st.lists(st.one_of(st.just(0), st.just(1), st.integers(2, 40), st.sampled_from([255, 256, 257, 511, 512, 513, 2047, 2048, 2049]),
                           st.integers(1000, 20000)), min_size=1, max_size=24)
TUNES = st.sampled_from([None, "vec=1,unroll=1,hint=0,block=128,ctas=2,item_tiles=1", "vec=2,unroll=2,hint=1,block=512,ctas=3,item_tiles=5",
                         "bulk=1,block=256,unroll=4,stages=2,ctas=2", "bulk=1,block=128,unroll=2,stages=3,ctas=4",
                         "bulk=1,block=256,unroll=2,stages=4,ctas=1"])


@settings(max_examples=40, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(sizes=SIZES, seed=st.integers(0, 2**31 - 1), tune=TUNES, mode=st.sampled_from([0, 1]), capacity=st.integers(64, 30000))
def test_random_batches_match_closed_form(capi, cuda, monkeypatch, sizes, seed, tune, mode, capacity):
    torch = cuda
    rng = np.random.default_rng(seed)
    n = int(sum(sizes))
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pts = helpers.synthetic_scan(max(n, 1), 64, seed % 1000)[:n]
    twists = [helpers.random_twist(rng) * rng.choice([0.0, 1.0, 1.0, 3.0]) for _ in sizes]
    x_reqs = rng.uniform(0, 1, len(sizes))
    params = capi.params_array([capi.frame_params_from_twist(x, float(r)) for x, r in zip(twists, x_reqs)])
    fracs = None
    if mode == capi.TIME_FROM_W:
        fracs = rng.uniform(0, 1, n).astype(np.float32)
        pts = pts.copy()
        pts[:, 3] = fracs
    if tune is None:
        monkeypatch.delenv("KMC_B200_TUNE", raising=False)
    else:
        monkeypatch.setenv("KMC_B200_TUNE", tune)
    guard = 32
    d_in = torch.from_numpy(pts).cuda() if n else torch.empty((0, 4), dtype=torch.float32, device="cuda")
    d_out = torch.full((n + guard, 4), -77.0, dtype=torch.float32, device="cuda")
    d_off = torch.from_numpy(offsets).cuda()
    d_par = torch.from_numpy(params.view(np.uint8)).cuda()
    capi.deskew_batch_device(d_in.data_ptr() if n else 0, d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(), len(sizes), n, mode,
                             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    assert np.all(out[n:] == -77.0), "wrote outside the batch"
    out = out[:n]
    for f, (xi, xr) in enumerate(zip(twists, x_reqs)):
        a, b = offsets[f], offsets[f + 1]
        if a == b:
            continue
        want = helpers.closed_form_deskew(pts[a:b], xi, float(np.float32(xr)), None if fracs is None else fracs[a:b].astype(np.float64))
        # accuracy model of DESIGN.md §3: float32 output rounding + ~2.5e-7 of the displacement (the x3 twists reach 50 m)
        disp = float(np.abs(want - pts[a:b, :3].astype(np.float64)).max())
        assert np.abs(out[a:b, :3] - want).max() < max(TOL_M, 4e-6 + 3e-7 * disp), (f, sizes)
    assert out[:, 3].tobytes() == pts[:, 3].tobytes()
    if n:
        with capi.Handle(0, capacity) as h:
            assert h.deskew_batch(pts, offsets, params, mode=mode).tobytes() == out.tobytes()


@settings(max_examples=30, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(sizes=SIZES, seed=st.integers(0, 2**31 - 1), capacity=st.sampled_from([64, 1000, 4096, 20_000, 250_000]),
       memory=st.sampled_from(["pageable", "pinned", "pinned+16"]), zero_copy=st.booleans(), in_place=st.booleans(),
       frame_call=st.booleans())
def test_random_host_calls_match_the_device_result(capi, cuda, monkeypatch, sizes, seed, capacity, memory, zero_copy, in_place, frame_call):
    """The host entry points over both transports (zero copy: the kernel reads / writes pinned memory itself; copy engines: the
    three-slot H2D / kernel / D2H pipeline), pageable and pinned caller memory (also pinned memory that is only 16-byte aligned,
    which switches the kernels to 128-bit accesses), in place and out of place, staging capacities far below and above the
    batch: always the device entry point's bits, and nothing written outside the batch."""
    torch = cuda
    rng = np.random.default_rng(seed)
    if frame_call:
        sizes = [int(sum(sizes))]
    n = int(sum(sizes))
    if n == 0:
        return
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    pts = helpers.synthetic_scan(n, 64, seed % 1000)
    params = capi.params_array([capi.frame_params_from_twist(helpers.random_twist(rng), float(rng.uniform(0, 1))) for _ in sizes])
    monkeypatch.setenv("KMC_B200_TUNE", "zc_points=100000000" if zero_copy else "zc_points=0")
    d_in = torch.from_numpy(pts).cuda()
    d_out = torch.empty_like(d_in)
    d_off = torch.from_numpy(offsets).cuda()  # named: a temporary would be freed (and its block reused) before the kernel runs
    d_par = torch.from_numpy(params.view(np.uint8)).cuda()
    capi.deskew_batch_device(d_in.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(), len(sizes), n, 0,
                             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = d_out.cpu().numpy()
    guard = 8
    shift = 1 if memory == "pinned+16" else 0

    def buffer(fill):
        if memory == "pageable":
            whole = np.full((n + guard + 1, 4), fill, dtype=np.float32)
            return whole, whole[shift:shift + n], whole.ctypes.data + 16 * shift
        whole = torch.full((n + guard + 1, 4), fill, dtype=torch.float32).pin_memory()
        return whole, whole.numpy()[shift:shift + n], whole.data_ptr() + 16 * shift

    keep_in, view_in, ptr_in = buffer(0.0)
    view_in[:] = pts
    if in_place:
        keep_out, view_out, ptr_out = keep_in, view_in, ptr_in
    else:
        keep_out, view_out, ptr_out = buffer(-77.0)
    with capi.Handle(0, capacity) as h:
        if frame_call:
            h.deskew_frame_ptr(ptr_in, ptr_out, n, capi.FrameParams.from_buffer_copy(params[:1].tobytes()))
        else:
            h.deskew_batch_ptr(ptr_in, ptr_out, offsets, params)
    assert view_out.tobytes() == want.tobytes(), (memory, zero_copy, in_place, capacity, sizes)
    whole_out = keep_out if isinstance(keep_out, np.ndarray) else keep_out.numpy()
    fill = 0.0 if in_place else -77.0
    assert np.all(whole_out[:shift] == fill) and np.all(whole_out[shift + n:] == fill), "wrote outside the caller's buffer"
