#!/usr/bin/env python
"""One launch of every secondary kernel (projection variants, reference-layout f64 deskew, single-frame deskew, pseudo
time stamps) on 20 M points, meant to run under `ncu --set full` so that their DRAM traffic per point can be put beside
their algorithmic bytes:

  ncu --set full --clock-control none -k regex:'Project|F64|DeskewFrame|PseudoTime|Checksums|CheckFractions' -o gpurun_out/secondary python tools/ncu_secondary.py
  python tools/ncu_secondary.py --summarize gpurun_out/secondary.ncu-rep profiles/r02_ncu_secondary_kernels.csv
"""
import csv
import io
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

N = 20_000_000
# kernel-name substring, template flags that tell the variants apart, algorithmic bytes per point
ALGORITHMIC = [
    ("DeskewFrameKernel", 32), ("PseudoTimeStampsXyKernel", 24), ("PseudoTimeStampsKernel", 24), ("DeskewCloudF64BatchKernel", 72),
    ("DeskewCloudF64Kernel", 72), ("FrameChecksumsKernel", 16), ("CheckFractionsKernel", 16),
    ("DeskewProjectBatchKernel<0, 1, 1", 48), ("DeskewProjectBatchKernel<0, 4, 0", 80), ("DeskewProjectBatchKernel<0, 4, 1", 96),
    ("ProjectFrameKernel<0, 0", 32), ("ProjectFrameKernel<1, 1", 48), ("ProjectFrameKernel<1, 0", 32),
    ("ProjectFrame4Kernel<0, 0", 80), ("ProjectFrame4Kernel<1, 1", 96),
]


def launch_all():
    import torch
    from kitti_motion_compensation_b200 import capi
    torch.cuda.set_device(0)
    s = torch.cuda.current_stream().cuda_stream
    d_in = torch.empty((N, 4), dtype=torch.float32, device="cuda")
    capi.synth_scans_device(d_in.data_ptr(), N, 1, 128, 20110926, 0, s)
    d_out = torch.empty_like(d_in)
    planes = [torch.empty_like(d_in) for _ in range(4)]
    T = np.eye(4)
    T[:3, :3] = np.array([7.533745e-03, -9.999714e-01, -6.166020e-04, 1.480249e-02, 7.280733e-04, -9.998902e-01, 9.998621e-01,
                          7.523790e-03, 1.480755e-02]).reshape(3, 3)
    T[:3, 3] = [-4.069766e-03, -7.631618e-02, -2.717806e-01]
    R_rect = np.array([9.999239e-01, 9.837760e-03, -7.445048e-03, -9.869795e-03, 9.999421e-01, -4.278459e-03, 7.402527e-03,
                       4.351614e-03, 9.999631e-01]).reshape(3, 3)
    P2 = np.array([7.215377e+02, 0, 6.095593e+02, 4.485728e+01, 0, 7.215377e+02, 1.728540e+02, 2.163791e-01, 0, 0, 1, 2.745884e-03]).reshape(3, 4)
    cam = capi.camera_params_from_calibration(P2, R_rect, T, 15.0)
    params, _ = capi.synth_frame_params(1, 20110926, 0, 0.5)
    p = capi.FrameParams.from_buffer_copy(params.tobytes())
    pp = [q.data_ptr() for q in planes]
    capi.deskew_frame_device(d_in.data_ptr(), d_out.data_ptr(), N, p, 0, s)
    capi.project_frame_device(d_in.data_ptr(), planes[0].data_ptr(), N, cam, s)
    capi.deskew_project_frame_device(d_in.data_ptr(), d_out.data_ptr(), planes[0].data_ptr(), N, p, cam, 0, s)
    capi.deskew_project_frame_device(d_in.data_ptr(), 0, planes[0].data_ptr(), N, p, cam, 0, s)
    capi.deskew_project_frame4_device(d_in.data_ptr(), 0, pp, N, None, [cam] * 4, 0, s)
    capi.deskew_project_frame4_device(d_in.data_ptr(), d_out.data_ptr(), pp, N, p, [cam] * 4, 0, s)
    # reference layout: column-major double cloud + stamps -> column-major double
    cloud = torch.empty((4, N), dtype=torch.float64, device="cuda")
    cloud[:3] = d_in[:, :3].T.double()
    cloud[3] = 1.0
    stamps = torch.rand(N, dtype=torch.float64, device="cuda") * 0.1
    out64 = torch.empty_like(cloud)
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    lib = capi.lib()
    import ctypes as C
    lib.kmc_b200_deskew_cloud_f64_device.restype = C.c_int
    rc = lib.kmc_b200_deskew_cloud_f64_device(C.c_void_p(cloud.data_ptr()), C.c_void_p(stamps.data_ptr()), C.c_void_p(out64.data_ptr()),
                                              C.c_int64(N), C.c_double(0.0), C.c_double(0.1), C.c_double(0.05), C.byref(p),
                                              C.c_void_p(flags.data_ptr()), C.c_void_p(s))
    assert rc == 0, capi.last_error()
    st = torch.empty(N, dtype=torch.float64, device="cuda")
    lib.kmc_b200_pseudo_time_stamps_device(C.c_void_p(d_in.data_ptr()), C.c_void_p(st.data_ptr()), C.c_int64(N), C.c_double(0.0),
                                           C.c_double(0.1), C.c_void_p(s))
    # GetPseudoTimeStamps on the reference's double columns
    lib.kmc_b200_pseudo_time_stamps_xy_device(C.c_void_p(cloud[0].data_ptr()), C.c_void_p(cloud[1].data_ptr()), C.c_void_p(st.data_ptr()),
                                              C.c_int64(N), C.c_double(0.0), C.c_double(0.1), C.c_void_p(s))
    # batch forms: 160 frames of 125 000 points = N
    F, per = 160, N // 160
    offs = torch.arange(0, (F + 1) * per, per, dtype=torch.int64, device="cuda")
    bparams, _ = capi.synth_frame_params(F, 20110926, 0, 0.5)
    d_par = torch.from_numpy(bparams.view(np.uint8).copy()).cuda()
    times = torch.tensor([[0.0, 0.1, 0.05]] * F, dtype=torch.float64, device="cuda").reshape(-1)
    bflags = torch.zeros(F, dtype=torch.int32, device="cuda")
    flat = cloud.reshape(-1)
    for f in range(F):  # frame f's w column inside its own 4 x per block
        flat[4 * f * per + 3 * per:4 * (f + 1) * per] = 1.0
    capi.deskew_cloud_f64_batch_device(flat.data_ptr(), stamps.data_ptr(), out64.data_ptr(), offs.data_ptr(), d_par.data_ptr(), times.data_ptr(), F, N,
                                       bflags.data_ptr(), s)
    capi.deskew_project_batch_device(d_in.data_ptr(), d_out.data_ptr(), pp[:1], offs.data_ptr(), d_par.data_ptr(), F, N, [cam], 0, s)
    capi.deskew_project_batch_device(d_in.data_ptr(), 0, pp, offs.data_ptr(), d_par.data_ptr(), F, N, [cam] * 4, 0, s)
    capi.deskew_project_batch_device(d_in.data_ptr(), d_out.data_ptr(), pp, offs.data_ptr(), d_par.data_ptr(), F, N, [cam] * 4, 0, s)
    sums = torch.zeros(F, dtype=torch.int64, device="cuda")
    capi.frame_checksums_device(d_out.data_ptr(), offs.data_ptr(), F, N, sums.data_ptr(), s)
    capi.check_fractions_device(d_in.data_ptr(), N, bflags.data_ptr(), s)
    torch.cuda.synchronize()


def summarize(rep, out_csv):
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                                                     check=True).stdout)))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}

    def val(r, name):
        i = hdr.index(name)
        return float(r[i].replace(',', '')) * scale.get(units[i], 1.0)

    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "block", "grid", "registers", "duration_us", "dram_read_bytes_per_point", "dram_write_bytes_per_point",
                    "dram_bytes_per_point", "algorithmic_bytes_per_point", "dram_throughput_pct_of_peak", "achieved_GBps_algorithmic"])
        for r in rows:
            name = r[hdr.index("Kernel Name")]
            alg = next((b for key, b in ALGORITHMIC if key in name), None)
            if alg is None:
                continue
            dur_i = hdr.index("gpu__time_duration.sum")
            dur = float(r[dur_i].replace(',', '')) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[dur_i], 1.0)  # -> us
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            w.writerow([name.split("(")[0], r[hdr.index("Block Size")], r[hdr.index("Grid Size")], r[hdr.index("launch__registers_per_thread")],
                        round(dur, 1), round(rd / N, 2), round(wr / N, 2), round((rd + wr) / N, 2), alg,
                        r[hdr.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")], round(alg * N / dur / 1e3, 0)])
    print(open(out_csv).read())


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "--summarize":
        summarize(sys.argv[2], sys.argv[3])
    else:
        launch_all()
