#!/bin/bash
# scan6 iteration pass: FindAll parity tests, c3 bench at several tunings, one full ncu capture of the scan.
tag=${1:-s6}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "find_all or fast_scan or forced" > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload c3 --steps 10 --no-e2e --no-cpu --no-parity > $out/bench_c3_$name.json 2> $out/bench_c3_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_c3_$name.json"))
    r=d["roofline"]; print("$name value %.0f GB/s  scan %.3f ms chain %.3f emit %.3f frac %.3f matches %d recs %d" % (d["value"], r["scan_ms"], r["chain_ms"], r["emit_ms"], r["frac"], d["run_info"]["matches_per_step"], d["run_info"]["distinct_records_per_step"]))
except Exception as e:
    print("$name failed", e); print(open("$out/bench_c3_$name.err").read()[-2000:])
PY
}
run G3PF0 RGX_SCAN_GROUPS=3 RGX_SCAN_PF=0
run G3PF1 RGX_SCAN_GROUPS=3 RGX_SCAN_PF=1
run G4PF1 RGX_SCAN_GROUPS=4 RGX_SCAN_PF=1
run G3PF1W48 RGX_SCAN_GROUPS=3 RGX_SCAN_PF=1 RGX_SCAN_WALK_AT=48
run G3PF1W64 RGX_SCAN_GROUPS=3 RGX_SCAN_PF=1 RGX_SCAN_WALK_AT=64
run G3PF1W16 RGX_SCAN_GROUPS=3 RGX_SCAN_PF=1 RGX_SCAN_WALK_AT=16
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:findall_scan6 -s 3 -c 1 -o $out/scan6_c3 \
   python bench.py --workload c3 --gib 1 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > $out/ncu_full_c3.log 2>&1
tail -2 $out/ncu_full_c3.log
fi
