#!/bin/bash
# 2-GPU A/B of the N2 launch variants; run under gpurun --gpus 2
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 5 --no-other --no-cpu"
for i in 1 2; do
  $T > gpurun_out/r02u_n2_base_$i.json 2> gpurun_out/r02u_n2_base_$i.err
  WF_VARIANT=0,0,0,8 $T > gpurun_out/r02u_n2_pdl_$i.json 2> gpurun_out/r02u_n2_pdl_$i.err
done
for f in gpurun_out/r02u_n2_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(j.get("value"), j.get("ms_per_step"), json.dumps(j.get("timeline"))[:400])
PY
done
