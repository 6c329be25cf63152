#!/bin/bash
out=gpurun_out/${1:-r2h}
mkdir -p $out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/pattern_table.py --only IPv4,TDFASemVer,URLCapture --mib 16 > $out/memcheck_set.log 2>&1; echo "rc=$?" >> $out/memcheck_set.log
grep -v "Host Frame\|^=========$" $out/memcheck_set.log | head -12 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
timeout 240 python tools/pattern_table.py --mib 512 > $out/pattern_table.jsonl 2> $out/pattern_table.err
cat $out/pattern_table.jsonl | cut -c1-420; tail -3 $out/pattern_table.err
timeout 300 python bench.py --workload c3 --steps 10 --no-e2e --no-cpu --no-parity > $out/bench_c3.json 2> $out/bench_c3.err; cut -c1-600 $out/bench_c3.json
