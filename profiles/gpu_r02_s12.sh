#!/bin/bash
# round 2, session 12: VD-CMA sampling kernel with a warp-private shared-memory row (wide rows), vd_update v3
tag=r02s12
mkdir -p gpurun_out
for f in test_gpu_es test_gpu_sizes test_gpu_parity test_gpu_l3; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
python profiles/vd_clocks.py > gpurun_out/${tag}_vd_clocks.txt 2>&1
cat gpurun_out/${tag}_vd_clocks.txt
python profiles/prof_cfg.py slopes > gpurun_out/${tag}_slopes.txt 2>&1
head -3 gpurun_out/${tag}_slopes.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_vd.csv \
   python profiles/prof_cfg.py vd > gpurun_out/${tag}_launches_vd.log 2>&1
python profiles/launch_summary.py gpurun_out/${tag}_launches_vd.csv | head -8
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"vd_sample" -s 2 -c 1 -f -o /tmp/${tag}_vd_sample \
     python profiles/prof_cfg.py vd > gpurun_out/${tag}_ncu_vd_sample.log 2>&1
python profiles/summarize_ncu.py /tmp/${tag}_vd_sample.ncu-rep gpurun_out/${tag}_vd_sample_ncu_summary.json 16384 >> gpurun_out/${tag}_ncu_vd_sample.log 2>&1
du -sh gpurun_out
