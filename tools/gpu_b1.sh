#!/bin/bash
# bench/infra pass: batch tests, all four bench workloads with in-run parity.
tag=${1:-b1}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "match_multi or corpus_batch or c1_date or forced" > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
show() {
python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}; c=d.get("cpu_baseline") or {}
    print("$2: value %.1f GB/s ms/step %.3f | kernel %s frac %.3f | e2e %s | cpu %s (1 core %s) | parity %s | launches %s" % (
        d["value"], d["ms_per_step"], r["kernel"], r["frac"], e.get("value"), c.get("value"), c.get("one_core_value"), d.get("parity"), d.get("gpu_launches")))
    print("   run_info", d.get("run_info"))
except Exception as ex:
    print("$2 failed", ex); print(open("$1".replace(".json",".err")).read()[-3000:])
PY
}
timeout 600 python bench.py --workload c4 --steps 10 > $out/bench_c4.json 2> $out/bench_c4.err; show $out/bench_c4.json c4
timeout 600 python bench.py --workload c3 --steps 10 > $out/bench_c3.json 2> $out/bench_c3.err; show $out/bench_c3.json c3
timeout 600 python bench.py --workload c2 --steps 10 > $out/bench_c2.json 2> $out/bench_c2.err; show $out/bench_c2.json c2
timeout 600 python bench.py --workload c5 --gib 2 --steps 5 > $out/bench_c5.json 2> $out/bench_c5.err; show $out/bench_c5.json c5
ls $out
