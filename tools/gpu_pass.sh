#!/bin/bash
# One GPU pass over the two FindAll workloads (the round-1 evidence sequence; tools/gpu_final.sh covers all four workloads):
# parity tests, both FindAll benches, ncu launch lists and one full capture of each scan kernel.
# Usage (under gpurun): bash tools/gpu_pass.sh <tag>
tag=${1:-pass}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $out/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -3 $out/pytest.log
timeout 600 python bench.py --workload c3 > $out/bench_c3.json 2> $out/bench_c3.err; tail -c 3000 $out/bench_c3.json
timeout 600 python bench.py --workload c2 > $out/bench_c2.json 2> $out/bench_c2.err; tail -c 3000 $out/bench_c2.json
kill $SMI
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'findall|rgx|scan|chain|emit' -c 200 --csv --log-file $out/launches_c3.csv \
   python bench.py --workload c3 --steps 3 --warmup 3 --no-e2e --no-cpu > $out/ncu_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'findall|rgx|scan|chain|emit' -c 200 --csv --log-file $out/launches_c2.csv \
   python bench.py --workload c2 --steps 3 --warmup 3 --no-e2e --no-cpu > $out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:findall_scan6 -s 3 -c 1 -o $out/scan6_c3 \
   python bench.py --workload c3 --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_full_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:findall_scan_btrun -s 3 -c 1 -o $out/btrun_c2 \
   python bench.py --workload c2 --steps 1 --warmup 3 --no-e2e --no-cpu > $out/ncu_full_c2.log 2>&1
fi
ls -la $out
