#!/bin/bash
# What the driver runs at round end, on one B200: smoke(), the GPU test suite, the default bench line, the reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-final}
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${T}_smoke.log 2>&1
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 ) > gpurun_out/${T}_tests.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench1.json 2> gpurun_out/${T}_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
( timeout 300 python examples/headless_driver.py 4 /tmp/hns_cache 2>&1 | tail -5 ) > gpurun_out/${T}_headless.log 2>&1
cat gpurun_out/${T}_smoke.log gpurun_out/${T}_tests.log gpurun_out/${T}_headless.log; cut -c1-330 gpurun_out/${T}_bench1.json; echo; cut -c1-400 gpurun_out/${T}_bench_ref.json; echo; cut -c1-300 gpurun_out/${T}_bench_c3.json; tail -2 gpurun_out/${T}_bench_ref.err
