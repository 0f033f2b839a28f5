"""Summarise ncu outputs into profiles/ (tracked): a launch list and the key counters of a full capture.

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches_c2.md "title"
    python tools/ncu_summary.py full gpurun_out/prof_bmu_tc.ncu-rep profiles/r1_bmu_tc_c2.md "title"
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_membar",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_bytes.sum",
]


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    seq = [(r["Kernel Name"].split("(")[0].replace("void ", ""), float(r["Metric Value"]) / 1e3) for r in rows]
    agg = collections.OrderedDict()
    for n, us in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# %s\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are "
                "cold-cache and serialised: compare SHARES).\n\n" % title)
        f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
        for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (n, c, us, us / c, 100 * us / tot))
        f.write("\n## launch sequence (first 40)\n\n```\n")
        for n, us in seq[:40]:
            f.write("%10.1f us  %s\n" % (us, n))
        f.write("```\n")


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (`%s`), one column per "
                "captured launch.\n\n| metric | unit | %s |\n|---|---|%s\n"
                % (title, src.split("/")[-1], " | ".join("launch %d" % i for i in range(len(data))),
                   "---:|" * len(data)))
        f.write("| kernel | | %s |\n" % " | ".join("`%s`" % r[idx["Kernel Name"]].split("(")[0] for r in data))
        for k in KEYS:
            if k in idx:
                f.write("| %s | %s | %s |\n" % (k, units[idx[k]], " | ".join(r[idx[k]] for r in data)))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else dst
    (launches if mode == "launches" else full)(src, dst, title)
