# One GPU call that refreshes everything under profiles/ (run through gpurun from the repo root):
#   tests, per-launch metrics of ONE serialised bench step, ncu --set full captures of the real-scale marches and of the
#   cascade, the launch list of the overlapped pipeline, and the bench line itself.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q) > gpurun_out/tests_r1d.log 2>&1; tail -2 gpurun_out/tests_r1d.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
ACFB_OVERLAP=0 timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/step_metrics_r1d.csv -s 84 -c 28 python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/nm.log 2>&1
ACFB_OVERLAP=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_smooth|k_trix|k_triyhist|k_gradmag" -s 64 -c 4 -o gpurun_out/prof_r1d_real -f python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/nf.log 2>&1
ACFB_OVERLAP=0 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_cascade" -s 12 -c 1 -o gpurun_out/prof_r1d_cascade -f python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline > gpurun_out/nf2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv -s 168 -c 56 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/nl.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -c 400 gpurun_out/bench_r1d.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1d_reference.json 2>> gpurun_out/bench_r1d.err; tail -c 300 gpurun_out/bench_r1d_reference.json
ls -la gpurun_out | tail -12
