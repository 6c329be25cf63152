#!/bin/bash
tag=${1:-c5}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_stream.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
for bs in 0 1048576; do
timeout 600 python bench.py --workload c5 --gib 2 --steps 5 --no-cpu --buffer-size $bs > $out/bench_c5_$bs.json 2> $out/bench_c5_$bs.err
python - <<PY
import json
try:
    d=json.loads(open("$out/bench_c5_$bs.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}
    print("c5 B=$bs: value %.1f GB/s ms/step %.3f | frac %.4f | e2e %s | parity %s | launches %s" % (d["value"], d["ms_per_step"], r["frac"], e.get("value"), d["parity"]["ok"], d["gpu_launches"]))
except Exception as ex:
    print("failed", ex); print(open("$out/bench_c5_$bs.err").read()[-3000:])
PY
done
if [ "$2" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'find_reader|exclusive' -c 60 --csv --log-file $out/launches_c5.csv \
   python bench.py --workload c5 --gib 2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > $out/ncu_c5.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$out/launches_c5.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:60], r[-1], "ns")
PY
fi
