"""Summarise ncu outputs into profiles/ (tracked).  Usage:
  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv  profiles/r1_launches.md
  python tools/ncu_summary.py full     gpurun_out/prof_r1.ncu-rep  profiles/r1_kernels.md
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")) / 1e6)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: `{src}`  command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n")
        f.write("| kernel | launches | mean ms | total ms | share |\n|---|---|---|---|---|\n")
        for k, v in agg.items():
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.3f} | {sum(v):.3f} | {100*sum(v)/tot:.1f} % |\n")
        f.write("\nper-launch ms in capture order:\n\n")
        for k, v in agg.items():
            f.write(f"* `{k}`: {', '.join(f'{x:.2f}' for x in v[:16])}\n")


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src}` (--clock-control none)\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[hdr.index('Kernel Name')]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for m, label in WANT:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {label} (`{m}`) | {r[i]} | {units[i]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
