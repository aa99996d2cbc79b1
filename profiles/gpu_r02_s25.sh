#!/bin/bash
# round 2, session 25: the build with vdcma split into per-dtype translation units (clean rebuild: 2m20 instead of 8m30)
tag=r02s25
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/${tag}_pytest_gpu.log; tail -1 gpurun_out/${tag}_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ) > gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_smoke.log
python profiles/vd_clocks.py 2>&1 | grep -E "total|timeline"
python profiles/prof_cfg.py slopes 2>&1 | head -2
