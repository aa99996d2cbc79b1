#!/bin/bash
# round 2, session 2: VD-CMA rework (fused sampling constants, factored weighted sums, fused reduce+update)
tag=r02s2
mkdir -p gpurun_out
rm -f gpurun_out/size_parity.jsonl gpurun_out/l3_stats.jsonl
( timeout 900 python -m pytest tests/test_gpu_sizes.py tests/test_gpu_l3.py tests/test_gpu_es.py -m gpu -q -x 2>&1 | tail -150 ) > gpurun_out/${tag}_pytest_es.log
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -80 ) > gpurun_out/${tag}_pytest_all.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_vd.csv \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_launches_vd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vd_sample_eval|vd_wsum|vd_update" -s 3 -c 3 -f -o gpurun_out/${tag}_vd \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_ncu_vd.log 2>&1
timeout 600 python profiles/prof_cfg.py slopes > gpurun_out/${tag}_slopes.txt 2>&1
tail -4 gpurun_out/${tag}_pytest_es.log; tail -4 gpurun_out/${tag}_pytest_all.log; cat gpurun_out/${tag}_slopes.txt
