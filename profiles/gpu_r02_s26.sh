#!/bin/bash
# round 2, session 26 (2 GPUs): seeds on concurrent streams (test + bench c5_seeds), final 2-GPU bench line
tag=r02s26
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parallel.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${tag}_pytest_parallel.log; tail -2 gpurun_out/${tag}_pytest_parallel.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_2gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "value_l2_resident")}, d["e2e"]["value"], d["config"]["cold_passes_ms"])
for k in ("c5_seeds", "extras_error"):
    print(k, d.get(k))
print({k: d["c3_sharded"][k] for k in ("us_per_gen_peer", "us_per_gen_nccl", "us_per_gen_1gpu", "bitwise_equal_to_1gpu")})
PY
tail -3 gpurun_out/${tag}_bench_2gpu.err
