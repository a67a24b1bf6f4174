#!/usr/bin/env python
"""Condense an `ncu -i x.ncu-rep --page raw --csv` export into the JSON summary kept under profiles/.
usage: python tools/ncu_summary.py raw.csv out.json [note]"""
import csv, json, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        s = {"kernel": d.get("Kernel Name", "")}
        for k in KEYS:
            if k in d and d[k] != "":
                s[k] = [d[k], u.get(k, "")]
        st = [(float(v), h) for h, v in d.items()
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v]
        tot = sum(x for x, _ in st) or 1.0
        s["stall_breakdown_pct"] = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: round(100 * x / tot, 1)
                                    for x, h in sorted(st, reverse=True)[:8]}
        out.append(s)
    res = {"source": "ncu --set full --clock-control none (one launch), exported with --page raw --csv",
           "note": sys.argv[3] if len(sys.argv) > 3 else "", "launches": out}
    json.dump(res, open(sys.argv[2], "w"), indent=1)
    for s in out:
        print(s["kernel"][:100], s.get("gpu__time_duration.sum"), s["stall_breakdown_pct"])


if __name__ == "__main__":
    main()
