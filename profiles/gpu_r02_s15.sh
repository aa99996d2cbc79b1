#!/bin/bash
# round 2, session 15 (2 GPUs): VD-CMA with z of the next generation drawn inside the update kernel; where a
# generation of the row-sharded CPSO goes (stage ablation)
tag=r02s15
mkdir -p gpurun_out
for f in test_gpu_es test_gpu_sizes test_gpu_l3; do
  ( timeout 1200 python -m pytest tests/$f.py -m gpu -q 2>&1 | tail -40 ) > gpurun_out/${tag}_pytest_$f.log
  echo "$f: $(tail -1 gpurun_out/${tag}_pytest_$f.log)"
done
python profiles/vd_clocks.py > gpurun_out/${tag}_vd_clocks.txt 2>&1
grep -E "total|timeline" gpurun_out/${tag}_vd_clocks.txt
python profiles/prof_cfg.py slopes > gpurun_out/${tag}_slopes.txt 2>&1
cat gpurun_out/${tag}_slopes.txt
SP_VD_NO_ZGEN=1 python profiles/prof_cfg.py slopes 2>&1 | head -2 > gpurun_out/${tag}_slopes_no_zgen.txt
cat gpurun_out/${tag}_slopes_no_zgen.txt
python profiles/prof_cfg.py eigh_time 2>&1 | grep "N=256" > gpurun_out/${tag}_eigh_time.txt; cat gpurun_out/${tag}_eigh_time.txt
for sk in 0 8 12 14 15; do
  SP_SHARD_EAGER=1 SP_SHARD_SKIP=$sk timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
     --master-port 29519 profiles/prof_sharded_stages.py 2>/dev/null | grep "us/gen" >> gpurun_out/${tag}_sharded_stages.txt
done
for sk in 0 8 12 14 15; do
  SP_SHARD_EAGER=1 SP_SHARD_SKIP=$sk timeout 300 python profiles/prof_sharded_stages.py 2>/dev/null | grep "us/gen" >> gpurun_out/${tag}_sharded_stages.txt
done
cat gpurun_out/${tag}_sharded_stages.txt
