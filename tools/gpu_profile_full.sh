# ncu --set full captures of the octave-0 launch of each hot kernel in the 4th bench step (run through gpurun from the repo root):
#   bash tools/gpu_profile_full.sh <tag> [kernels...]   -> gpurun_out/<tag>_<kernel>.ncu-rep
# The skip counts are launches of the SAME kernel before the one captured (3 warm-up steps x launches per step).
TAG=${1:-r2}; shift
KS=${@:-"k_cascade_tail k_cascade_tile k_front k_triyhist k_chan"}
mkdir -p gpurun_out
for K in $KS; do
  case $K in k_chan) SKIP=15;; k_color|k_post) SKIP=3;; *) SKIP=12;; esac
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"^$K" -s $SKIP -c 1 -f -o gpurun_out/${TAG}_$K \
      python bench.py --steps 1 --warmup 3 --batch 256 --no-cpu-baseline --no-other-configs --no-parity > gpurun_out/${TAG}_$K.log 2>&1
  ls -la gpurun_out/${TAG}_$K.ncu-rep
done
