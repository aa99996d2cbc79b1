#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list and one full capture of the DE kernel.
# Usage (from repo root, on the GPU box): bash profiles/gpu_session.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${tag}_launches_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:de_pool_kernel -s 10 -c 2 -f -o gpurun_out/${tag}_de_pool \
   python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_run.log 2>&1
timeout 900 python bench_configs.py --quick > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_bench.json
