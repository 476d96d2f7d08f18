#!/bin/bash
# A/B timing of KKT kernel variants on one box: tools/ab_kkt.sh <out.log> <lib1> <lib2> ...   ("-" = the in-tree libdto.so)
OUT=$1; shift
for lib in "$@"; do
  for c in "cartpole 4096" "acrobot 4096" "acrobot 64"; do
    set -- $c
    if [ "$lib" = "-" ]; then unset DTO_LIB; else export DTO_LIB=$PWD/$lib; fi
    python tools/run_kkt.py --model $1 --T 101 --batch $2 --steps 40 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['model'], d['B'], round(d['ms_kkt_kernels'],4))" >> $OUT
  done
done
