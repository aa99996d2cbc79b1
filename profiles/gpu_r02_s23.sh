#!/bin/bash
# round 2, session 23 (2 GPUs): sharded CPSO with the bound decision riding the best exchange (one exchange and one
# launch per generation in the quiet phase)
tag=r02s23
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parallel.py tests/test_gpu_parity.py -m gpu -q -k "parallel or cpso or restart or pso or shard" 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_2gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "value_l2_resident")}, d["e2e"]["value"])
for k in ("c3_sharded", "c3_sharded_p262144", "c5_seeds", "extras_error"):
    print(k, d.get(k))
PY
tail -3 gpurun_out/${tag}_bench_2gpu.err
