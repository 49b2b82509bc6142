"""Key metrics of every kernel instance in an .ncu-rep (read on the dev box: `ncu -i rep --page raw --csv`) -> JSON list.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/ncu_x.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__cluster_dim_x", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    hi = rows.index(hdr)
    units = rows[hi + 1]
    res = []
    for r in rows[hi + 2:]:
        if len(r) != len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")][:160]}
        for k in KEYS:
            if k in hdr:
                v = r[hdr.index(k)].replace(",", "")
                try:
                    d[k] = float(v)
                except ValueError:
                    d[k] = v
                d.setdefault("_units", {})[k] = units[hdr.index(k)]
        res.append(d)
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
