#!/bin/bash
# round 2, session 31: the final library once more (smoke, quick parity subset) and bench.py with NO flags
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 )
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sizes.py -m gpu -x -q 2>&1 | tail -1 )
( time timeout 900 python bench.py > gpurun_out/r02s31_bench_default_flags.json 2> gpurun_out/r02s31_bench_default_flags.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s31_bench_default_flags.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "steps", "warmup", "ms_per_step", "value_l2_resident")}, d["config"]["cold_passes_ms"])
print(d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_generation"], 1) for k, v in d.get("configs", {}).items()}, d.get("extras_error"))
print(d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline", {}).get("kind"), d["clocks"])
PY
