#!/bin/bash
# ncu --set full capture of one kernel of the native solver:  tools/gpu_profile_resolve.sh <tag> <kernel regex> <B> [skip]
set -u
TAG=$1; KERNEL=$2; B=$3; SKIP=${4:-5}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$KERNEL" --launch-skip $SKIP -c 1 -f -o gpurun_out/$TAG \
    python tools/solve_config3.py $B 101 20 acrobot native > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_src.csv > gpurun_out/$TAG.txt 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
