#!/bin/bash
# round 2, session 22: CPSO restart decision inside the generation kernel (bound on the swarm radius), eager windows
# of 32 generations that end when the restarts stop
tag=r02s22
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_l3.py tests/test_parallel.py -m gpu -q 2>&1 | tail -30 ) > gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
SP_CPSO_EXACT_RADIUS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cpso or restart" 2>&1 | tail -2
python profiles/prof_timeline.py cpso 450 2>&1 | grep -v -i warn > gpurun_out/${tag}_cpso_timeline.txt; head -8 gpurun_out/${tag}_cpso_timeline.txt
SP_CPSO_EXACT_RADIUS=1 python profiles/prof_timeline.py cpso 450 2>&1 | grep -v -i warn > gpurun_out/${tag}_cpso_timeline_exact.txt; head -8 gpurun_out/${tag}_cpso_timeline_exact.txt
python profiles/prof_timeline.py cpso 200 2>&1 | grep -v -i warn | head -3
python profiles/prof_timeline.py pso 200 2>&1 | grep -v -i warn | head -2
