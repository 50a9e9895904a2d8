"""Summarise ncu outputs into profiles/ (tracked).  Usage:
  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv  profiles/r1_launches.md
  python tools/ncu_summary.py full     gpurun_out/prof_r1.ncu-rep  profiles/r1_kernels.md
  python tools/ncu_summary.py traffic  gpurun_out/prof_r1.ncu-rep  profiles/r1_traffic.json   (one bench step, kernels serialised)
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
        f.write(f"source: `{src}`  command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 60 python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n")
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


def traffic(src, dst):
    """DRAM read+write bytes and duration per kernel group of ONE captured bench step (bench.py reads this file for
    roofline.traffic).  The capture must hold exactly one step: ACFB_OVERLAP=0 ... -s 15 -c 15 ... --steps 1 --warmup 3."""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = {m: hdr.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    per = collections.OrderedDict()
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        g = per.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
        g["launches"] += 1
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            g["dram_bytes"] += float(r[cols[m]].replace(",", "")) * scale[units[cols[m]]]
        g["ms"] += float(r[cols["gpu__time_duration.sum"]].replace(",", "")) * scale[units[cols["gpu__time_duration.sum"]]]
    doc = {"source": f"{src} (ncu --set full --clock-control none, bench.py --steps 1 --warmup 3 --batch 256, kernels serialised with ACFB_OVERLAP=0)",
           "workload": {"rows": 1080, "cols": 1920, "model": "face80", "batch": 256, "operating_point": "fast"},
           "per_step": per}
    json.dump(doc, open(dst, "w"), indent=1)


def traffic_csv(src, dst):
    """Same as `traffic`, from the CSV log of an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
    smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --csv --log-file ...` pass over one bench step
    (a --set full report of all 33 launches of a step is too large to bring back)."""
    import json
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, ui, vi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "inst": 1.0, "%": 1.0}
    per = collections.OrderedDict(); seen = set()
    # keep ONE step: the launches from the second-last k_color launch up to (not including) the last one
    starts = sorted({int(r[ii]) for r in rows[1:] if r[ki].split("(")[0].replace("void ", "").split("<")[0] == "k_color"})
    lo, hi = (starts[-2], starts[-1]) if len(starts) >= 2 else (-1, 1 << 60)
    for r in rows[1:]:
        if not (lo <= int(r[ii]) < hi):
            continue
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        g = per.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0, "warp_instructions": 0.0, "issue_active_pct_time_weighted": 0.0})
        if (r[ii], name) not in seen:
            seen.add((r[ii], name)); g["launches"] += 1
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        if r[mi].startswith("dram__bytes"): g["dram_bytes"] += v
        elif r[mi].startswith("gpu__time_duration"): g["ms"] += v; g.setdefault("_t", []).append(v)
        elif r[mi].startswith("smsp__inst_executed"): g["warp_instructions"] += v
        elif r[mi].startswith("smsp__issue_active"): g.setdefault("_i", []).append(v)
    for g in per.values():
        t, i = g.pop("_t", []), g.pop("_i", [])
        g["issue_active_pct_time_weighted"] = sum(a * b for a, b in zip(t, i)) / max(1e-12, sum(t)) if len(t) == len(i) else None
    doc = {"source": f"{src} (ncu --metrics ... --clock-control none, bench.py --steps 1 --warmup 3 --batch 256; ONE step = the launches from one k_color to the next)",
           "workload": {"rows": 1080, "cols": 1920, "model": "face80", "batch": 256, "operating_point": "hits"},
           "per_step": per}
    json.dump(doc, open(dst, "w"), indent=1)


def step_table(src, dst):
    """Per-launch table of ONE bench step (second-last k_color launch up to the last one) from the CSV log of
    tools/gpu_profile_pass.sh: duration, DRAM read / write, L2 bytes, warp instructions, issue-active, warps active."""
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, mi, ui, vi, ii, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID", "Grid Size", "Block Size"))
    base = lambda r: r[ki].split("(")[0].replace("void ", "")
    starts = sorted({int(r[ii]) for r in rows[1:] if base(r).split("<")[0] == "k_color"})
    lo, hi = (starts[-2], starts[-1]) if len(starts) >= 2 else (-1, 1 << 60)
    L = collections.OrderedDict()
    for r in rows[1:]:
        if lo <= int(r[ii]) < hi:
            L.setdefault(int(r[ii]), {"k": base(r), "grid": r[gi], "block": r[bi]})[r[mi]] = float(r[vi].replace(",", ""))
    tot = sum(d["gpu__time_duration.sum"] for d in L.values()) / 1e6
    with open(dst, "w") as f:
        f.write("# One bench step, launch by launch (ncu --metrics ..., --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: `{src}`, command: `bash tools/gpu_profile_pass.sh` (bench.py --steps 1 --warmup 3 --batch 256, 1080p, face80, hits operating point, one pipeline)\n\n")
        f.write("| # | kernel | grid | block | ms | share | DRAM read GB | DRAM write GB | L2 GB | warp instr (M) | issue-active % | warps active % |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for n, (i, d) in enumerate(L.items()):
            ms = d["gpu__time_duration.sum"] / 1e6
            f.write(f"| {n} | `{d['k']}` | {d['grid']} | {d['block']} | {ms:.3f} | {100 * ms / tot:.1f} % | {d['dram__bytes_read.sum'] / 1e9:.2f} | {d['dram__bytes_write.sum'] / 1e9:.2f} | "
                    f"{d.get('lts__t_bytes.sum', 0) / 1e9:.2f} | {d['smsp__inst_executed.sum'] / 1e6:.0f} | {d['smsp__issue_active.avg.pct_of_peak_sustained_active']:.0f} | "
                    f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.0f} |\n")
        f.write(f"\ntotal {tot:.2f} ms over {len(L)} launches\n\n| kernel | launches | ms | share |\n|---|---|---|---|\n")
        agg = collections.OrderedDict()
        for d in L.values():
            a = agg.setdefault(d["k"].split("<")[0], [0, 0.0]); a[0] += 1; a[1] += d["gpu__time_duration.sum"] / 1e6
        for k, (n, ms) in agg.items():
            f.write(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f} % |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "traffic_csv": traffic_csv, "step_table": step_table}[sys.argv[1]](sys.argv[2], sys.argv[3])
