#!/bin/bash
# Quick GPU check: parity tests then a short device-only bench of one workload.   bash tools/gpu_quick.sh <tag> [c3|c2] [ncu-kernel-regex]
tag=${1:-q}; wl=${2:-c3}; kre=$3
out=gpurun_out/$tag
mkdir -p $out
timeout 240 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
for w in $wl; do
timeout 150 python bench.py --workload $w --no-e2e --no-cpu --steps 10 > $out/bench_$w.json 2> $out/bench_$w.err; tail -c 2500 $out/bench_$w.json; tail -5 $out/bench_$w.err
done
if [ -n "$kre" ]; then
w=${wl%% *}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kre -s 3 -c 1 -o $out/full_$w \
   python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_full_$w.log 2>&1
tail -3 $out/ncu_full_$w.log
fi
