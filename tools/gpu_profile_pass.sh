mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
ACFB_OVERLAP=0 timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/step_metrics_r1c.csv -s 84 -c 28 python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/nm.log 2>&1
ACFB_OVERLAP=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_smooth|k_trix|k_triyhist" -s 48 -c 3 -o gpurun_out/prof_r1c_real -f python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/nf.log 2>&1
ACFB_OVERLAP=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_cascade|k_chan" -s 36 -c 3 -o gpurun_out/prof_r1c_det -f python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/nf2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv -s 168 -c 56 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nl.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -c 600 gpurun_out/bench_r1c.json
ls -la gpurun_out
