"""The path on MORE THAN ONE GPU of a box (reference analogue: the serial frame loop of
handlers.cpp:55-64, whose frames are independent).  Frames are sharded in contiguous ranges, there is no collective, and a
frame's result must not depend on which GPU — or how many — produced it.

Skips below two visible devices; run with `gpurun --gpus 8 -- python -m pytest tests/test_multi_gpu.py -m gpu -v`
(the log of such a run is committed under profiles/).
"""
import numpy as np
import pytest

import helpers
from oracle import ref_binding as rb
from test_deskew_gpu import assert_parity, batch_params, make_batch, run_batch

pytestmark = pytest.mark.gpu


def n_devices(capi) -> int:
    return min(capi.lib().kmc_b200_device_count(), 8)


@pytest.fixture
def multi(capi, cuda):
    n = n_devices(capi)
    if n < 2:
        pytest.skip(f"needs at least 2 CUDA devices, this box shows {n}")
    return n


def test_batch_multi_gpu_matches_one_device_bitwise_and_the_reference(capi, oracle, cuda, multi):
    """kmc_b200_deskew_batch_multi_gpu on every visible device: bit-identical to one device, parity with the oracle and
    with the reference's own compiled sources on every frame; ragged frames, empty frames, fewer frames than devices."""
    n_dev = multi
    sizes = [25_000 + 977 * k for k in range(3 * n_dev + 1)]
    sizes[1] = 0
    sizes[n_dev] = 1
    sizes[-2] = 130_000
    pts, offsets, frames = make_batch(oracle, sizes, 8800, mercator=True)
    params = batch_params(capi, frames)
    want = run_batch(cuda, capi, pts, offsets, params)  # one launch on device 0
    handles = [capi.Handle(d, 40_000) for d in range(n_dev)]  # capacity below most frames: chunks cut frames on every device
    try:
        assert sorted(h.device for h in handles) == list(range(n_dev))
        for subset in (handles, handles[:2], handles[::-1]):
            got = capi.deskew_batch_multi_gpu(subset, pts, offsets, params)
            assert got.tobytes() == want.tobytes(), f"{len(subset)} devices"
        # pinned caller memory (no staging), the layout bench.py's e2e_inprocess leg uses
        pin_in = cuda.from_numpy(pts).pin_memory()
        pin_out = cuda.zeros_like(pin_in).pin_memory()
        import ctypes as C
        arr = (C.c_void_p * n_dev)(*[h.raw for h in handles])
        capi.check(capi.lib().kmc_b200_deskew_batch_multi_gpu(arr, n_dev, pin_in.data_ptr(), pin_out.data_ptr(), offsets.ctypes.data,
                                                              params.ctypes.data, len(sizes), capi.TIME_FROM_AZIMUTH))
        assert pin_out.numpy().tobytes() == want.tobytes()
        # fewer frames than devices: trailing devices get nothing
        few = capi.deskew_batch_multi_gpu(handles, pts[: offsets[1]], offsets[:2], params[:1])
        assert few.tobytes() == want[: offsets[1]].tobytes()
        # two handles on the same device are refused
        with capi.Handle(0, 1000) as dup:
            with pytest.raises(capi.KmcError) as e:
                capi.deskew_batch_multi_gpu([handles[0], dup], pts, offsets, params)
            assert e.value.status == capi.ERR_BAD_SIZE
    finally:
        for h in handles:
            h.close()
    for f, (Ts, Te, _, xr) in enumerate(frames):
        a, b = offsets[f], offsets[f + 1]
        if a == b:
            continue
        t_req = 10.0 + xr * 0.1
        assert_parity(want[a:b], oracle.deskew_xyzi_scan(pts[a:b], Ts, Te, 10.0, 10.1, t_req), pts[a:b])
        if rb.available():
            assert_parity(want[a:b], rb.deskew_xyzi_scan(pts[a:b], Ts, Te, 10.0, 10.1, t_req), pts[a:b])


def test_every_device_computes_the_same_bits_and_checksums(capi, oracle, cuda, multi):
    """The same batch through the DEVICE entry point on each GPU in turn: identical output bits and identical per-frame
    checksums (what bench.py's strong-scaling leg compares across ranks), and the device checksum equals its definition."""
    torch = cuda
    sizes = [130_000, 0, 99_999, 3, 64_001]
    pts, offsets, frames = make_batch(oracle, sizes, 8900)
    params = batch_params(capi, frames)
    outs, sums = [], []
    try:
        for d in range(multi):
            torch.cuda.set_device(d)
            dev = torch.device("cuda", d)
            d_in = torch.from_numpy(pts).to(dev)
            d_out = torch.empty_like(d_in)
            d_off = torch.from_numpy(offsets).to(dev)
            d_par = torch.from_numpy(params.view(np.uint8)).to(dev)
            d_sum = torch.empty(len(sizes), dtype=torch.int64, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            capi.deskew_batch_device(d_in.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_par.data_ptr(), len(sizes), len(pts), 0, st)
            capi.frame_checksums_device(d_out.data_ptr(), d_off.data_ptr(), len(sizes), len(pts), d_sum.data_ptr(), st)
            torch.cuda.synchronize(dev)
            outs.append(d_out.cpu().numpy())
            sums.append(d_sum.cpu().numpy().view(np.uint64))
    finally:
        torch.cuda.set_device(0)
    for d in range(1, multi):
        assert outs[d].tobytes() == outs[0].tobytes(), f"device {d}"
        assert np.array_equal(sums[d], sums[0]), f"device {d}"
    assert np.array_equal(sums[0], capi.frame_checksums_numpy(outs[0], offsets))
    assert sums[0][1] == 0 and len(set(sums[0].tolist())) == len(sizes)


def test_drop_in_library_on_another_device(capi, cuda, multi):
    """Handles live on any device: the reference-layout (double) host call on the LAST device gives device 0's bits."""
    pts = helpers.real_scan()[:50_000]
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    p = capi.frame_params_from_poses(T_start, T_end, t0, t2, t1)
    cloud = np.concatenate([pts[:, :3].astype(np.float64), np.ones((len(pts), 1))], axis=1)
    from oracle import binding as ob
    stamps = ob.pseudo_time_stamps(cloud, t0, t2)
    res = []
    for d in (0, multi - 1):
        with capi.Handle(d, 250_000) as h:
            out, flags, rc = h.deskew_cloud_f64(cloud, stamps, t0, t2, t1, p)
            assert rc == capi.OK and flags == 0
            res.append(out)
    assert res[0].tobytes() == res[1].tobytes()
