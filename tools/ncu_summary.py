#!/usr/bin/env python
"""Turns an `ncu --set full` report of the deskew kernel into the summaries committed under profiles/:
  <prefix>_ncu_full_key_metrics.csv, <prefix>_ncu_full_details.csv, <prefix>_ncu_stalls.csv and profiles/roofline_traffic.json.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01 <points per launch>
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = ['ID', 'Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes.sum.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__cycles_elapsed.avg.per_second', 'sm__cycles_elapsed.avg', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, prefix, points = sys.argv[1], sys.argv[2], int(sys.argv[3])
    raw = page(rep, "raw")
    hdr, units, launches = raw[0], raw[1], raw[2:]
    with open(prefix + "_ncu_full_key_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        for i, h in enumerate(hdr):
            if h in KEEP:
                w.writerow([h, units[i]] + [r[i] for r in launches])
    with open(prefix + "_ncu_stalls.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["warps_stalled_per_issue_active", "launch0"])
        rows = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    rows.append((float(launches[0][i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        for v, h in sorted(rows, reverse=True):
            w.writerow([h, v])
    with open(prefix + "_ncu_full_details.csv", "w", newline="") as f:
        csv.writer(f).writerows(page(rep, "details"))
    r = launches[0]
    rd = float(r[hdr.index('dram__bytes_read.sum')].replace(',', '')) * SCALE[units[hdr.index('dram__bytes_read.sum')]]
    wr = float(r[hdr.index('dram__bytes_write.sum')].replace(',', '')) * SCALE[units[hdr.index('dram__bytes_write.sum')]]
    traffic = {"source": os.path.basename(prefix) + "_ncu_full_key_metrics.csv (ncu --set full, launch 0: " + r[hdr.index('Kernel Name')].split('(')[0].strip() + ")",
               "points": points, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_point": (rd + wr) / points,
               "algorithmic_bytes_per_point": 32}
    with open(os.path.join(os.path.dirname(prefix) or ".", "roofline_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
