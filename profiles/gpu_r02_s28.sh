#!/bin/bash
# round 2, session 28 (8 GPUs): the driver's bench command at N = 8 with the final code
tag=r02s28_8gpu
mkdir -p gpurun_out
nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "value_l2_resident")}, d["e2e"]["value"], d["config"]["cold_passes_ms"])
for k in ("c3_sharded", "c3_sharded_p262144", "c5_seeds", "extras_error"):
    print(k, d.get(k))
PY
tail -3 gpurun_out/${tag}_bench.err
