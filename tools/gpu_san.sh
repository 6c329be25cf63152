#!/bin/bash
# compute-sanitizer over smoke() (memcheck, then racecheck on the FindAll part), and the fuzz tests
out=gpurun_out/${1:-san}
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q > $out/fuzz.log 2>&1; echo "fuzz rc=$?" >> $out/fuzz.log; tail -5 $out/fuzz.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $out/memcheck.log 2>&1; echo "memcheck rc=$?" >> $out/memcheck.log
tail -6 $out/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "
import regengo_b200 as rg
from regengo_b200 import synth
p = rg.Pattern(synth.URL_PATTERN)
print(p.find_all_offsets(synth.make_buffer('url', 600000))[0])
p2 = rg.Pattern(synth.EMAIL_PATTERN)
print(p2.find_all_offsets(synth.make_buffer('log', 300000))[0])
" > $out/racecheck.log 2>&1; echo "racecheck rc=$?" >> $out/racecheck.log
tail -6 $out/racecheck.log
