#!/bin/bash
# Evidence pass of a round: bench lines of the four workloads, ncu launch lists of the same commands, one full ncu
# capture per dominant kernel.  bash tools/gpu_final.sh <tag>
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
for w in c3 c2 c4 c5; do
  timeout 600 python bench.py --workload $w > $out/bench_$w.json 2> $out/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/bench_$w.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}; c=d.get("cpu_baseline") or {}
    print("$w: value %.1f %s  ms/step %.3f | kernel %s frac %.4f | e2e %s | cpu %s (1 core %s) | parity %s | launches %s | clocks %s" % (
        d["value"], d["unit"], d["ms_per_step"], r.get("kernel"), r["frac"], e.get("value"), c.get("value"), c.get("one_core_value"), (d.get("parity") or {}).get("ok"), d["gpu_launches"], d.get("clocks")))
except Exception as ex:
    print("$w failed", ex); print(open("$out/bench_$w.err").read()[-2000:])
PY
done
for w in c3 c2 c4 c5; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"findall|rgx|match_multi|find_reader|replace_batch|exclusive_scan|max_len" -c 400 --csv --log-file $out/launches_$w.csv \
     python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu --no-parity > $out/ncu_$w.log 2>&1
done
cap() {  # name workload kernel-regex extra-args
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -o $out/full_$1 \
     python bench.py --workload $2 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity $4 > $out/ncu_full_$1.log 2>&1
  ls -la $out/full_$1.ncu-rep 2>/dev/null | awk '{print $5, $9}'
}
cap scan6_c3 c3 findall_scan6 "--gib 2"
cap emit_c3 c3 findall_emit3 "--gib 2"
cap btrun_c2 c2 findall_scan_btrun ""
cap multi_c4 c4 match_multi_kernel "--inputs 4000000"
cap chase_c5 c5 find_reader_chase "--gib 2"
cap records_c5 c5 find_reader_records "--gib 2"
ls -la $out
