"""Turns `ncu -i <rep> --page raw --csv` into the markdown summary kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv
    python tools/ncu_summary.py gpurun_out/prof_raw.csv "title" > profiles/<name>.md

One table per kernel (template arguments kept), averaged over its captured launches.
"""
import csv
import re
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
              "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0,
              "msecond": 1e3, "second": 1e6}


def short(name):
    m = re.match(r"(?:void )?([A-Za-z0-9_:]+(?:<[^(]*>)?)", name)
    return m.group(1) if m else name


def main():
    path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu summary"
    rows = list(csv.reader(open(path)))
    # header row, units row, then one row per launch
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    header, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {}
    for i, h in enumerate(header):
        for k in KEYS:
            if h == k or h.endswith("." + k):
                col.setdefault(k, i)
    kcol = header.index("Kernel Name")
    gcol, bcol = header.index("Grid Size"), header.index("Block Size")
    per = OrderedDict()
    for r in data:
        if len(r) <= kcol:
            continue
        per.setdefault((short(r[kcol]), r[gcol], r[bcol]), []).append(r)
    print("# %s\n" % title)
    for (name, grid, block), rs in per.items():
        print("## `%s`  grid %s block %s  (%d launches captured)\n" % (name, grid, block, len(rs)))
        print("| metric | mean over launches | unit |\n|---|---|---|")
        for k in KEYS:
            if k not in col:
                continue
            vals = []
            for r in rs:
                try:
                    vals.append(float(r[col[k]].replace(",", "")))
                except ValueError:
                    pass
            if not vals:
                continue
            print("| %s | %.6g | %s |" % (k, sum(vals) / len(vals), units[col[k]]))
        # derived: DRAM bytes per launch and GB/s
        try:
            rd = [float(r[col["dram__bytes_read.sum"]].replace(",", "")) *
                  UNIT_SCALE.get(units[col["dram__bytes_read.sum"]], 1.0) for r in rs]
            wr = [float(r[col["dram__bytes_write.sum"]].replace(",", "")) *
                  UNIT_SCALE.get(units[col["dram__bytes_write.sum"]], 1.0) for r in rs]
            tm = [float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
                  UNIT_SCALE.get(units[col["gpu__time_duration.sum"]], 1.0) for r in rs]
            tot = (sum(rd) + sum(wr)) / len(rs)
            t = sum(tm) / len(tm)
            print("| **derived: DRAM bytes per launch (read + write)** | %.6g | byte |" % tot)
            print("| **derived: DRAM GB/s under ncu (cold, serialised)** | %.5g | GB/s |" % (tot / t / 1e3))
        except (KeyError, ValueError, ZeroDivisionError):
            pass
        print()


if __name__ == "__main__":
    main()
