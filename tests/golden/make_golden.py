"""Writes tests/golden/oracle_real_scan_digest.json from the CPU oracle (config 1 of BASELINE.json / SURVEY 8d)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import binding as ob  # noqa: E402
import helpers  # noqa: E402


def main():
    pts = helpers.real_scan()
    T_start, T_end, t0, t1, t2 = helpers.config1_frame()
    ref = ob.deskew_xyzi_scan(pts, T_start, T_end, t0, t2, t1)
    idx = np.linspace(0, len(pts) - 1, 64).astype(int)
    digest = {
        "_comment": "Oracle (CPU restatement) output for the real scan under the config-1 motion; see README.md.",
        "delta_twist": helpers.CONFIG1_TWIST,
        "sample_index": idx.tolist(),
        "sample_xyz": ref[idx, :3].tolist(),
        "sum_xyz": ref[:, :3].sum(axis=0).tolist(),
        "sum_abs_displacement": float(np.abs(ref[:, :3] - pts[:, :3].astype(np.float64)).sum()),
        "max_displacement": float(np.abs(ref[:, :3] - pts[:, :3].astype(np.float64)).max()),
    }
    with open(os.path.join(HERE, "oracle_real_scan_digest.json"), "w") as f:
        json.dump(digest, f, indent=1)
    print("max displacement", digest["max_displacement"])


if __name__ == "__main__":
    main()
