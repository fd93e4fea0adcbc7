#!/usr/bin/env python
"""Turns ncu exports into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/<launches>.csv  > profiles/<name>_launches.md
  python profiles/summarize.py raw      gpurun_out/<raw>.csv       > profiles/<name>_ncu.md
(the raw csv comes from `ncu -i <rep> --page raw --csv`)."""
import csv
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "launch__cluster_size", "sm__cycles_active.avg",
]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        try:
            ns = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[kn]
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ns; a[2] = max(a[2], ns)
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share | max ms |\n|---|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda t: -t[1][1]):
        print(f"| `{name[:110]}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e6:.3f} |")
    print(f"\ntotal kernel time {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches "
          "(ncu-serialised, cold cache: compare shares, not absolutes)")


def raw(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"### `{r[hdr.index('Kernel Name')][:120]}`  (launch id {r[0]})\n")
        print("| metric | unit | value |\n|---|---|---:|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {units[i]} | {r[i]} |")
        st = []
        for i, h in enumerate(hdr):
            if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    st.append((float(r[i].replace(",", "")), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("\nwarp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6]) + "\n")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
