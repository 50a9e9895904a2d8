# One GPU call that refreshes the per-launch table of ONE bench step (run through gpurun from the repo root):
#   bash tools/gpu_profile_pass.sh <tag>      -> gpurun_out/<tag>_step_metrics.csv (dram bytes, duration, instructions, issue-active per launch)
# The ncu --set full captures (profiles/r2_kernels_*.md) come from tools/gpu_profile_full.sh.
TAG=${1:-r2}
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum
# warm-up 3 steps + other launches come first: skip to the last timed step (-s), capture one step's launches (-c)
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_step_metrics.csv ${NCU_SKIP:+-s $NCU_SKIP} -c ${NCU_COUNT:-40} \
    python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline --no-other-configs --no-parity > gpurun_out/${TAG}_nm.log 2>&1
tail -c 400 gpurun_out/${TAG}_nm.log
