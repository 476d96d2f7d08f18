#!/bin/bash
# ncu --set full capture of the banded LDL' factor kernel:  tools/gpu_profile_kkt.sh <tag> <model> <T> <B>
set -u
TAG=$1; MODEL=$2; T=$3; B=$4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:kkt_band_kernel" --launch-skip 3 -c 1 -f -o gpurun_out/$TAG \
    python tools/run_kkt.py --model $MODEL --T $T --batch $B --steps 3 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_src.csv > gpurun_out/$TAG.txt 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
