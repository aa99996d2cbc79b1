#!/bin/bash
# round 2, session 3: VD-CMA iteration 2 (PDL chain, 256-thread update, keyed Philox), CMA-ES panel gather + pipelined GEMMs
tag=r02s3
mkdir -p gpurun_out
rm -f gpurun_out/size_parity.jsonl gpurun_out/l3_stats.jsonl
( timeout 900 python -m pytest tests/test_gpu_sizes.py tests/test_gpu_es.py -m gpu -q 2>&1 | tail -150 ) > gpurun_out/${tag}_pytest_es.log
( timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_sizes.py --deselect tests/test_gpu_es.py 2>&1 | tail -80 ) > gpurun_out/${tag}_pytest_rest.log
for w in vd cma; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_${w}.csv \
   python profiles/prof_cfg.py $w > gpurun_out/${tag}_launches_${w}.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vd_sample_eval|vd_wsum|vd_update" -s 3 -c 3 -f -o gpurun_out/${tag}_vd \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_ncu_vd.log 2>&1
timeout 600 python profiles/prof_cfg.py slopes > gpurun_out/${tag}_slopes.txt 2>&1
tail -4 gpurun_out/${tag}_pytest_es.log; tail -4 gpurun_out/${tag}_pytest_rest.log; cat gpurun_out/${tag}_slopes.txt
