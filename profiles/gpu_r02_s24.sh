#!/bin/bash
# round 2, session 24 (2 GPUs): final validation -- whole GPU suite, smoke, 1-GPU bench + reference arm, 2-GPU bench
tag=r02s24
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/${tag}_pytest_gpu.log; tail -2 gpurun_out/${tag}_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ) > gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_smoke.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s24_bench_1gpu.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "value_dirty_flush", "value_l2_resident", "value_kernel_only")}, d["config"]["cold_passes_ms"])
print(d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_generation"], 1) for k, v in d.get("configs", {}).items()})
print(d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"), d.get("extras_error"))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_2gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "value_l2_resident")}, d["e2e"]["value"], d["config"]["cold_passes_ms"])
for k in ("c3_sharded", "c3_sharded_p262144", "c5_seeds", "extras_error"):
    print(k, d.get(k))
PY
