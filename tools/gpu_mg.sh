#!/bin/bash
# multi-GPU pass (run under gpurun --gpus N): NCCL parity tests, then the sharded benches
N=${1:-2}
tag=${2:-mg}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -15 $out/pytest.log
for wl in c3 c4 c5; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $wl --steps 10 --warmup 3 --no-cpu > $out/bench_${wl}_n$N.json 2> $out/bench_${wl}_n$N.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$out/bench_${wl}_n$N.json") if l.startswith("{")][-1]
    e=d.get("e2e") or {}
    print("$wl N=$N: value %.1f GB/s ms/step %.3f | e2e %.1f (%.1f ms) | parity %s | %s" % (d["value"], d["ms_per_step"], e.get("value") or 0, e.get("ms_per_step") or 0, d["parity"]["ok"], d["run_info"].get("sharding")))
except Exception as ex:
    print("$wl failed", ex); print(open("$out/bench_${wl}_n$N.err").read()[-3000:])
PY
done
