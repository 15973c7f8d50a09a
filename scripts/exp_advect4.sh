#!/bin/bash
# packed (third-generation) advection: full GPU suite, per-stage times with the switch on and off, launch list, ncu of the new kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 ) > gpurun_out/exp6_tests.log 2>&1
HNS_ADVECT4=1 timeout 300 python scripts/time_phases_gpu.py c4 8 > gpurun_out/exp6_phases_on.txt 2>&1
HNS_ADVECT4=0 timeout 300 python scripts/time_phases_gpu.py c4 8 > gpurun_out/exp6_phases_off.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/exp6_bench.json 2> gpurun_out/exp6_bench.err
cat gpurun_out/exp6_tests.log gpurun_out/exp6_phases_on.txt gpurun_out/exp6_phases_off.txt; cut -c1-600 gpurun_out/exp6_bench.json; tail -3 gpurun_out/exp6_bench.err
