#!/bin/bash
# round 2, session 20: merge half of the ranking fused into vd_wsum (small populations)
tag=r02s20
mkdir -p gpurun_out
for f in test_gpu_es test_gpu_sizes test_gpu_l3 test_gpu_parity; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
python profiles/vd_clocks.py > gpurun_out/${tag}_vd_clocks.txt 2>&1
grep -E "total|timeline" gpurun_out/${tag}_vd_clocks.txt
python profiles/prof_cfg.py slopes 2>&1 | head -2 > gpurun_out/${tag}_slopes.txt; cat gpurun_out/${tag}_slopes.txt
SP_VD_NO_FUSED_RANK=1 python profiles/prof_cfg.py slopes 2>&1 | head -2 > gpurun_out/${tag}_slopes_unfused.txt; cat gpurun_out/${tag}_slopes_unfused.txt
