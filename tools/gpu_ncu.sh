#!/bin/bash
# one full ncu capture:  bash tools/gpu_ncu.sh <tag> <workload> <kernel-regex> <skip>
tag=$1; w=$2; kre=$3; skip=${4:-3}
out=gpurun_out/$tag; mkdir -p $out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -o $out/full_$w \
   python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_full_$w.log 2>&1
tail -2 $out/ncu_full_$w.log | cut -c1-200
