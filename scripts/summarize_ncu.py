#!/usr/bin/env python
"""Turns ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python scripts/summarize_ncu.py kernel   gpurun_out/prof_forward_r1.ncu-rep profiles/r1_forward_kernel.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as fh:
        fh.write("# Launch list (ncu --metrics gpu__time_duration.sum --clock-control none): `%s`\n\n" % src)
        fh.write("Per-launch times are cold-cache and serialised; compare SHARES.\n\n| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k.split("(")[0][:70], n, t, t / n, 100 * t / tot))
    try:
        print(open(dst).read())
    except BrokenPipeError:
        pass


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = dict(zip(hdr, zip(vals, units)))
    stalls = []
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(d[k][0]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    with open(dst, "w") as fh:
        fh.write("# ncu --set full --clock-control none: `%s`\n\nkernel: `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % (src, d["Kernel Name"][0]))
        for k in KEYS:
            if k in d:
                fh.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
        fh.write("\nWarp stall reasons (warps per issue-active cycle):\n\n| reason | ratio |\n|---|---:|\n")
        for v, k in sorted(stalls, reverse=True)[:12]:
            fh.write("| %s | %.3f |\n" % (k, v))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
