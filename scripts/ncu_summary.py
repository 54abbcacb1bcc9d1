#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` export into the per-launch figures the roofline discussion uses.

    python scripts/ncu_summary.py gpurun_out/r02d_ncu_gemm_raw.csv > profiles/r02d_ncu_gemm_summary.json
"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_inst_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "sm__cycles_elapsed.avg.per_second": "sm_clock",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        rec = {"kernel": re.sub(r"\(.*", "", d["Kernel Name"]).replace("nerf::<unnamed>::", "").replace("nerf::(anonymous namespace)::", ""),
               "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for i, h in enumerate(hdr):
            base = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1] in ("TriageCompute",) else h
            for k, short in WANT.items():
                if h == k or h.endswith("." + k) or base == k:
                    try:
                        rec[short] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                    except ValueError:
                        pass
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
