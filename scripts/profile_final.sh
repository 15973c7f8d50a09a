#!/bin/bash
# End-of-round evidence on one B200: launch list of two consecutive frames (the second one runs every third-generation kernel),
# `ncu --set full` of the kernels that matter, the default bench line. Outputs under gpurun_out/ (copied to profiles/ by hand).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2e}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python scripts/profile_frame.py c4 2 > gpurun_out/${T}_ncu0.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launch_summary.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'advect_.*4|k_subtract_gradient|k_combustion_buoyancy_packed' -c 7 -o gpurun_out/${T}_packed -f python scripts/profile_frame.py c4 2 > gpurun_out/${T}_ncu1.log 2>&1
python scripts/ncu_summary.py gpurun_out/${T}_packed.ncu-rep > gpurun_out/${T}_packed_ncu.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rbgs_split' --launch-skip 100 -c 4 -o gpurun_out/${T}_rbgs -f python scripts/profile_frame.py c4 2 > gpurun_out/${T}_ncu2.log 2>&1
python scripts/ncu_summary.py gpurun_out/${T}_rbgs.ncu-rep --all > gpurun_out/${T}_rbgs_split_ncu.txt 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench1.json 2> gpurun_out/${T}_bench1.err
cat gpurun_out/${T}_launch_summary.txt; grep -E "===|time_duration|dram__bytes|lsu_wavefronts.avg" gpurun_out/${T}_packed_ncu.txt gpurun_out/${T}_rbgs_split_ncu.txt | cut -c1-150; cut -c1-300 gpurun_out/${T}_bench1.json; tail -2 gpurun_out/${T}_bench1.err
