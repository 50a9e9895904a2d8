"""python tools/ncu_hot.py report.ncu-rep [top_n] : key metrics + the source lines with most stall samples (needs -lineinfo, --import-source on)."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, r = rows[0], rows[2]
d = dict(zip(hdr, r))
def g(k):
    return d.get(k, "n/a")
print(g("Kernel Name")[:60], "grid", g("launch__grid_size"), "block", g("launch__block_size"), "regs", g("launch__registers_per_thread"),
      "smem/blk", g("launch__shared_mem_per_block_dynamic"), "occ_limit(warps)", g("launch__occupancy_limit_warps"), "theor.occ%", g("sm__maximum_warps_per_active_cycle_pct"))
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
          "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]:
    print(f"  {k} = {g(k)}")
st = [(k, float(v)) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
print("  stalls/issue:", ", ".join(f"{k[34:-23]}={v:.2f}" for k, v in sorted(st, key=lambda x: -x[1])[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r))
h = rows[hi]
ci = h.index("Source"); si = next(i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
ei = next((i for i, c in enumerate(h) if c == "Instructions Executed"), None)
body = [r for r in rows[hi + 1:] if len(r) > max(ci, si)]
tot = sum(float(r[si] or 0) for r in body); toti = sum(float(r[ei] or 0) for r in body) if ei is not None else 0
print(f"SASS lines {len(body)}, samples {tot:.0f}, inst {toti:.0f}")
agg = collections.Counter(); aggi = collections.Counter()
for r in body:
    op = r[ci].split()[0] if r[ci].split() else "?"
    if op.startswith("@"): op = r[ci].split()[1]
    agg[op.split(".")[0]] += float(r[si] or 0)
    if ei is not None: aggi[op.split(".")[0]] += float(r[ei] or 0)
print("by opcode (samples%):", ", ".join(f"{k}={100*v/tot:.1f}" for k, v in agg.most_common(12)))
if toti: print("by opcode (inst%):", ", ".join(f"{k}={100*v/toti:.1f}" for k, v in aggi.most_common(14)))
idx = sorted(range(len(body)), key=lambda i: -float(body[i][si] or 0))[:topn]
for i in sorted(idx):
    print(f"  [{i:4d}] {100*float(body[i][si] or 0)/tot:5.1f}%  {body[i][ci][:90]}")
