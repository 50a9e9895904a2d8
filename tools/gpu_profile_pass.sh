# One GPU call that refreshes profiles/ (run through gpurun from the repo root): per-launch metrics of ONE serialised bench
# step (30 launches: ACFB_OVERLAP=0, one lane), the launch list of the overlapped pipeline (2 lanes x 30), the bench lines.
# The ncu --set full captures (profiles/r1_kernels_*.md) come from:
#   ACFB_OVERLAP=0 ncu --set full --import-source on --clock-control none -k regex:"k_smooth|k_trix|k_triyhist|k_gradmag" -s 64 -c 4 -o gpurun_out/prof_real python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline
#   ACFB_OVERLAP=0 ncu --set full --import-source on --clock-control none -k regex:"k_cascade" -s 12 -c 1 -o gpurun_out/prof_cascade python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
ACFB_OVERLAP=0 timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/step_metrics_r1e.csv -s 90 -c 30 python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/nm.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv -s 180 -c 60 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nl.log 2>&1
python tools/ncu_summary.py traffic_csv gpurun_out/step_metrics_r1e.csv profiles/r1_traffic.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; tail -c 300 gpurun_out/bench_r1e.json
cp profiles/r1_traffic.json gpurun_out/r1_traffic_e.json
