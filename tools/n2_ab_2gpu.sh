#!/bin/bash
# 2-GPU A/B of launch variants of the distributed step (WF_VARIANT = E1,N1,E2,N2 variant numbers); run under gpurun --gpus 2.
# N1 variant 7 = k_halo_finish as an ordinary launch, N2 variant 8 = phase 4 as an ordinary launch (the default launches
# both as programmatic dependent launches; at the time of the runs logged in profiles/r02_scaling.md the numbers meant
# the opposite: 7 / 8 switched the dependent launches ON).
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 5 --no-other --no-cpu"
for i in 1 2; do
  for v in 0,0,0,0 0,7,0,0 0,0,0,8 0,7,0,8; do
    WF_VARIANT=$v $T > gpurun_out/ab2_${v//,/}_$i.json 2> gpurun_out/ab2_${v//,/}_$i.err
  done
done
for f in gpurun_out/ab2_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
t=j.get("timeline") or {}
print(j.get("value"), j.get("ms_per_step"), [round(v,4) for v in t.values() if isinstance(v,float)])
PY
done
