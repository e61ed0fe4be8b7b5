#!/usr/bin/env python
"""Summarise ncu output into small text files under profiles/ (the tracked evidence).

    python tools/ncu_summary.py launches gpurun_out/launches_X.csv profiles/X_launches.md
    python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep profiles/X_kernels.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaheavy%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_cyc%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_cyc%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wf%"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "lsu_wb%"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__inst_executed.sum", "inst"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("l3d::", "")


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): per-kernel totals\n\n")
        f.write("source: %s; cold-cache serialised launches -- compare SHARES, not absolutes\n\n" % src)
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f | %.1f%% |\n" % (k, n, t, t / n, 100 * t / tot))
        f.write("\ntotal %.1f us over %d launches\n" % (tot, sum(n for n, _ in agg.values())))


def full(src, dst, per_kernel=2):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, s) for m, s in METRICS if m in idx]
    seen = collections.Counter()
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none: key metrics per kernel (first %d launches each)\n\n" % per_kernel)
        f.write("source: %s\n\n" % src)
        f.write("| kernel | " + " | ".join(s for _, s in cols) + " |\n|---|" + "---:|" * len(cols) + "\n")
        for r in rows[2:]:
            name = short(r[idx["Kernel Name"]])
            seen[name] += 1
            if seen[name] > per_kernel:
                continue
            vals = []
            for m, _ in cols:
                v, u = r[idx[m]], units[idx[m]]
                try:
                    x = float(v.replace(",", ""))
                    v = ("%.3f" % x).rstrip("0").rstrip(".")
                except ValueError:
                    pass
                vals.append(v + (" " + u if u and u not in ("%", "inst") else ""))
            f.write("| %s | " % name + " | ".join(vals) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
