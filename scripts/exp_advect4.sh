#!/bin/bash
# packed (third-generation) advection incl. the sharded frame: full GPU suite, per-stage times, bench, ncu of the final kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/exp7_tests.log 2>&1
HNS_ADVECT4=1 timeout 300 python scripts/time_phases_gpu.py c4 8 > gpurun_out/exp7_phases_on.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/exp7_bench.json 2> gpurun_out/exp7_bench.err
cat gpurun_out/exp7_tests.log gpurun_out/exp7_phases_on.txt; cut -c1-400 gpurun_out/exp7_bench.json; tail -3 gpurun_out/exp7_bench.err
