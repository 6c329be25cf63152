#!/bin/bash
# the whole GPU test suite (bounded), then the c5 bench
out=gpurun_out/${1:-full}
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -14 $out/pytest.log
timeout 600 python bench.py --workload c5 --gib 2 --steps 5 --no-cpu > $out/bench_c5.json 2> $out/bench_c5.err
python - <<PY
import json
d=json.loads(open("$out/bench_c5.json").read().strip().splitlines()[-1])
print("c5: value %.1f GB/s ms/step %.3f parity %s launches %s" % (d["value"], d["ms_per_step"], d["parity"]["ok"], d["gpu_launches"]))
PY
