#!/bin/bash
# Run under gpurun (1 GPU). Collects: launch list, full ncu captures of the four step kernels.
# usage: tools/profile_gpu.sh <tag> [preload]
TAG=${1:-r01}
PRE=${2:-0}
OUT=gpurun_out
mkdir -p $OUT
# warm-up = 3 one-step calls (each N1 carries the folded first predictor), then ONE batch of 6 steps (steady-state launches)
BENCH="python bench.py --steps 6 --warmup 3 --no-cpu --no-other --no-sustained --preload $PRE"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(elem|node|predict)' -s $((PRE*4)) -c 40 --csv --log-file $OUT/launches_${TAG}.csv $BENCH > $OUT/launches_${TAG}.log 2>&1
for K in k_elem_main k_node_update k_elem_vol k_node_vol; do
  SKIP=$((PRE+4))
  if [ $K = k_node_vol ]; then SKIP=$((PRE+7)); fi   # launches 0-1 init, 2-4 warm-up, 5 first of the batch (with predictor), 6.. steady state
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $OUT/prof_${TAG}_$K $BENCH > $OUT/prof_${TAG}_$K.log 2>&1
done
ls -la $OUT
