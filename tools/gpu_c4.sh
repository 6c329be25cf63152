#!/bin/bash
tag=${1:-c4}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "match_multi or corpus_batch or c1_date or forced" > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 600 python bench.py --workload c4 --steps 10 --no-cpu > $out/bench_c4.json 2> $out/bench_c4.err
python - <<PY
import json
try:
    d=json.loads(open("$out/bench_c4.json").read().strip().splitlines()[-1])
    r=d["roofline"]; e=d.get("e2e") or {}
    print("c4: value %.1f GB/s ms/step %.3f | frac %.4f | e2e %s | parity %s | inputs/s %.3g" % (d["value"], d["ms_per_step"], r["frac"], e.get("value"), d["parity"]["ok"], d["run_info"]["inputs_per_s"]))
except Exception as ex:
    print("failed", ex); print(open("$out/bench_c4.err").read()[-3000:])
PY
if [ "$2" == "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_multi -s 3 -c 1 -o $out/mm_c4 \
   python bench.py --workload c4 --inputs 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
fi
