#!/bin/bash
# One ncu --set full capture of the dominant kernel of a workload + raw/source CSV pages (run under gpurun).
#   tools/gpu_profile.sh <tag> <model spec for tools/run_fused.py> [kernel regex]
# Outputs gpurun_out/<tag>.ncu-rep, <tag>_raw.csv, <tag>_src.csv, <tag>.txt (tools/ncu_summary.py)
set -u
TAG=$1; SPEC=$2; KRE=${3:-knot_kernel_ws}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$KRE" --launch-skip 4 -c 1 -f -o gpurun_out/$TAG \
    python tools/run_fused.py "$SPEC" 8 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_src.csv > gpurun_out/$TAG.txt 2>&1
tail -5 gpurun_out/${TAG}_ncu.log
