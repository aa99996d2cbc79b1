#!/bin/bash
# round 2, session 30: the undecided-bound test hook; whole GPU suite on the final library
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r02s30_pytest_gpu.log; tail -1 gpurun_out/r02s30_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 )
