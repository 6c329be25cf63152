#!/bin/bash
# ncu launch list of one workload's bench step:  bash tools/gpu_launches.sh <tag> <workload>
tag=${1:-l}; w=${2:-c3}
out=gpurun_out/$tag; mkdir -p $out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'findall|rgx|scan|chain|emit|batch|reader' -c 300 --csv --log-file $out/launches_$w.csv \
   python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu > $out/ncu_$w.log 2>&1
tail -2 $out/ncu_$w.log | cut -c1-300
