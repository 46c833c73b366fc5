"""kmc_b200_run_prepare — the host half of MotionCompensateRun (handlers.cpp:41-65 minus the point clouds): frame count,
time stamps, OxTS packets, MakeFrame poses, per-frame kernel records.  No GPU involved, so everything here runs on CPU:
the records are compared with records built from the ORACLE's MakeFrame poses and, when oracle/_ref is built, from the
poses of the reference's own compiled MakeFrame."""
import os
import shutil

import numpy as np
import pytest

import helpers
from oracle import ref_binding as rb


def _records_from_poses(capi, make_frame_poses, info, n):
    packets = [[info["middles"][i], *info["oxts"][i]] for i in range(n)]
    out = []
    for i in range(1, n - 1):
        T_start, T_end = make_frame_poses(packets[i - 1], packets[i], packets[i + 1], info["starts"][i], info["ends"][i])
        out.append(capi.frame_params_from_poses(T_start, T_end, info["starts"][i], info["ends"][i], info["middles"][i]))
    return capi.params_array(out)


def _assert_records_close(a, b):
    assert a.shape == b.shape
    fa = a.view(np.float32).reshape(len(a), -1)
    fb = b.view(np.float32).reshape(len(b), -1)
    # 9 decimals of the clock strings and %.13g of the packets are all that separates the two routes
    assert np.allclose(fa, fb, rtol=2e-5, atol=2e-7), np.abs(fa - fb).max()


def test_records_match_the_oracle_and_the_reference_make_frame(capi, oracle, tmp_path):
    n = 9
    run = tmp_path / "run_sync"
    info = helpers.make_run_folder(str(run), n, 50, seed=12)
    n_frames, params = capi.run_prepare(str(run))
    assert n_frames == n and params.shape == (n - 2,)
    _assert_records_close(params, _records_from_poses(capi, oracle.make_frame_poses, info, n))
    if rb.available():
        _assert_records_close(params, _records_from_poses(capi, rb.make_frame_poses, info, n))
    # the requested time is the camera trigger: x_req = (middle - start) / (end - start)
    want = np.array([(info["middles"][i] - info["starts"][i]) / (info["ends"][i] - info["starts"][i]) for i in range(1, n - 1)])
    assert np.allclose(params["x_req"], want, atol=1e-6)
    # the run moves: 11 m/s and 0.3 rad/s over a ~0.103 s scan
    rec = params.view(np.float32).reshape(n - 2, -1)
    rho = np.sqrt((rec[:, 4:7] ** 2).sum(axis=1) + (rec[:, 8:11] ** 2).sum(axis=1))  # |rho_perp + rho_par|
    assert np.all(rho > 0.2) and np.all(rho < 3.0), rho


def test_short_runs_and_counting(capi, tmp_path):
    for n in (1, 2, 3):
        run = tmp_path / f"run{n}"
        helpers.make_run_folder(str(run), n, 10, seed=n)
        n_frames, params = capi.run_prepare(str(run))
        assert n_frames == n and params.size == max(n - 2, 0)
    # every directory entry counts as a frame (handlers.cpp:15-17), so a stray file makes the stamp files too short
    (tmp_path / "run3" / "velodyne_points" / "data" / "notes.txt").write_text("x")
    with pytest.raises(capi.KmcError) as e:
        capi.run_prepare(str(tmp_path / "run3"))
    assert e.value.status == capi.ERR_IO


def test_error_statuses(capi, tmp_path):
    n = 6
    run = tmp_path / "run_sync"
    info = helpers.make_run_folder(str(run), n, 10, seed=3)
    import ctypes as C
    count = C.c_int64()
    small = np.zeros(2, dtype=capi.FRAME_PARAMS_DTYPE)
    assert capi.lib().kmc_b200_run_prepare(os.fsencode(str(run)), 2, small.ctypes.data, C.byref(count)) == capi.ERR_CAPACITY
    assert capi.lib().kmc_b200_run_prepare(None, 0, None, C.byref(count)) == capi.ERR_NULL_POINTER
    with pytest.raises(capi.KmcError) as e:
        capi.run_prepare(str(tmp_path / "nowhere"))
    assert e.value.status == capi.ERR_IO and "not a KITTI run folder" in str(e.value)

    def broken(change):
        work = tmp_path / "broken"
        if work.exists():
            shutil.rmtree(work)
        shutil.copytree(run, work)
        change(work)
        with pytest.raises(capi.KmcError) as err:
            capi.run_prepare(str(work))
        return err.value

    err = broken(lambda w: os.remove(w / "oxts" / "data" / f"{2:010d}.txt"))
    assert err.status == capi.ERR_IO and "Oxts" in str(err)
    err = broken(lambda w: os.remove(w / "velodyne_points" / "timestamps_end.txt"))
    assert err.status == capi.ERR_IO and "failed to open timestamp file" in str(err)
    err = broken(lambda w: (w / "oxts" / "data" / f"{4:010d}.txt").write_text("49.0 8.4 garbage\n"))
    assert err.status == capi.ERR_IO and "malformed Oxts packet" in str(err)

    def shorten(w):
        lines = (w / "velodyne_points" / "timestamps.txt").read_text().splitlines()
        (w / "velodyne_points" / "timestamps.txt").write_text("\n".join(lines[:-1]) + "\n")
    assert broken(shorten).status == capi.ERR_IO

    def garble(w):
        lines = (w / "velodyne_points" / "timestamps_start.txt").read_text().splitlines()
        lines[1] = "2011-09-26 not-a-clock"
        (w / "velodyne_points" / "timestamps_start.txt").write_text("\n".join(lines) + "\n")
    assert broken(garble).status == capi.ERR_IO

    def out_of_range(w):  # a scan that starts before the previous OxTS packet: the reference asserts here
        lines = (w / "velodyne_points" / "timestamps_start.txt").read_text().splitlines()
        lines[3] = helpers._clock(info["middles"][2] - 0.25)
        (w / "velodyne_points" / "timestamps_start.txt").write_text("\n".join(lines) + "\n")
    err = broken(out_of_range)
    assert err.status == capi.ERR_TIME_OUT_OF_RANGE and "frame 3" in str(err)


def test_corrupt_run_folders_give_a_status_never_a_crash(capi, tmp_path):
    """The reference parses the run's text files with std::stoi / stod / getline and throws or aborts on garbage
    (data_io.cpp:18-99).  kmc_b200_run_prepare must return: OK with finite records, or a negative status — never crash, hang
    or hand back NaN records — whatever one of the text files holds.  Random byte strings, truncations, line-level edits and
    number-level edits of each of the five kinds of text file, drawn by hypothesis (deterministic)."""
    import ctypes as C

    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    n = 6
    run = tmp_path / "run_sync"
    helpers.make_run_folder(str(run), n, 10, seed=21)
    targets = ["velodyne_points/timestamps_start.txt", "velodyne_points/timestamps.txt", "velodyne_points/timestamps_end.txt",
               "oxts/timestamps.txt", "oxts/data/0000000002.txt"]
    originals = {t: (run / t).read_bytes() for t in targets}
    tokens = st.sampled_from([b"nan", b"inf", b"-inf", b"1e999", b"-", b"", b" ", b"\x00", b"25:61:61.5", b"ab:cd:12.3", b"13:04", b"::",
                              b"9" * 400, b"0x10", b"1,5", b"\xff\xfe", b"2011-09-26 13:04:32.123456789", b"2011-09-26"])

    def mutate(data: bytes, kind: int, pos: float, token: bytes, blob: bytes) -> bytes:
        lines = data.split(b"\n")
        k = min(int(pos * len(lines)), len(lines) - 1)
        if kind == 0:
            return blob  # arbitrary bytes
        if kind == 1:
            return data[: int(pos * len(data))]  # truncated
        if kind == 2:
            lines[k] = token  # one line replaced
        elif kind == 3:
            fields = lines[k].split(b" ")
            fields[min(int(pos * 7919) % max(len(fields), 1), len(fields) - 1)] = token  # one field replaced
            lines[k] = b" ".join(fields)
        elif kind == 4:
            del lines[k]  # one line missing
        else:
            lines.insert(k, token)  # one line too many
        return b"\n".join(lines)

    params = np.zeros(n, dtype=capi.FRAME_PARAMS_DTYPE)
    count = C.c_int64()
    seen = set()

    @settings(max_examples=400, deadline=None, database=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(which=st.integers(0, len(targets) - 1), kind=st.integers(0, 5), pos=st.floats(0, 1, exclude_max=True), token=tokens,
           blob=st.binary(max_size=300))
    def run_one(which, kind, pos, token, blob):
        target = targets[which]
        (run / target).write_bytes(mutate(originals[target], kind, pos, token, blob))
        try:
            count.value = -1
            rc = capi.lib().kmc_b200_run_prepare(os.fsencode(str(run)), n, params.ctypes.data, C.byref(count))
        finally:
            (run / target).write_bytes(originals[target])
        seen.add(rc)
        assert rc <= 1, rc  # OK, a warning, or an error status
        if rc >= 0:
            assert count.value == n
            rec = params[: n - 2].view(np.float32)
            assert np.all(np.isfinite(rec)), (target, kind, token)
        else:
            assert capi.last_error() != ""

    run_one()
    assert capi.OK in seen and any(rc < 0 for rc in seen)  # both outcomes were exercised
    n_frames, _ = capi.run_prepare(str(run))  # the restored folder still parses
    assert n_frames == n
