#!/bin/bash
# round 2, session 19 (8 GPUs): the driver's multi-GPU bench command at N = 8 (headline + c3_sharded + c5_seeds),
# the sharded-swarm pytest on real GPUs
tag=r02s19_8gpu
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "value_l2_resident")}, d["e2e"]["value"])
for k in ("c3_sharded", "c3_sharded_p262144", "c5_seeds", "extras_error"):
    print(k, d.get(k))
PY
tail -3 gpurun_out/${tag}_bench.err
( timeout 600 python -m pytest tests/test_parallel.py -m gpu -q 2>&1 | tail -5 ) > gpurun_out/${tag}_pytest_parallel.log; tail -2 gpurun_out/${tag}_pytest_parallel.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02s19_4gpu_bench.json 2> gpurun_out/r02s19_4gpu_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02s19_4gpu_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "value_l2_resident")}, d["e2e"]["value"])
for k in ("c3_sharded", "c3_sharded_p262144", "c5_seeds", "extras_error"):
    print(k, d.get(k))
PY
