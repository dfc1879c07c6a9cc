#!/bin/bash
# Run under gpurun (1 GPU). Collects: launch list, full ncu captures of the four step kernels.
# usage: tools/profile_gpu.sh <tag> [preload]
TAG=${1:-r01}
PRE=${2:-0}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu --no-other --no-sustained --preload $PRE"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(elem|node|predict)' -s $((PRE*4)) -c 40 --csv --log-file $OUT/launches_${TAG}.csv $BENCH > $OUT/launches_${TAG}.log 2>&1
for K in k_elem_main k_node_update k_elem_vol k_node_vol; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s $((PRE+4)) -c 1 -f -o $OUT/prof_${TAG}_$K $BENCH > $OUT/prof_${TAG}_$K.log 2>&1
done
ls -la $OUT
