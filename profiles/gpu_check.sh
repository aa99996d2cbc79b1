#!/bin/bash
# quick GPU check: parity tests + config throughput + launch lists:  bash profiles/gpu_check.sh <tag> [pytest -k expr]
tag=${1:-chk}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tail -25 ) > gpurun_out/${tag}_pytest.log
timeout 900 python bench_configs.py --quick > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
for w in vd vd64 cpso; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_${w}.csv \
     python profiles/prof_cfg.py $w > gpurun_out/${tag}_launches_${w}.log 2>&1
done
cat gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_configs.jsonl | cut -c1-200
