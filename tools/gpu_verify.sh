#!/bin/bash
# final verification: the whole GPU suite, then racecheck over FindAll / FindReader / Replace on small inputs
out=gpurun_out/${1:-verify}
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log
tail -10 $out/pytest.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "
import regengo_b200 as rg
from regengo_b200 import synth
p = rg.Pattern(synth.URL_PATTERN)
print(p.find_all_offsets(synth.make_buffer('url', 600000))[0])
p2 = rg.Pattern(synth.EMAIL_PATTERN)
print(p2.find_all_offsets(synth.make_buffer('log', 300000))[0])
s = synth.make_buffer('stream', 300000, digit_noise=0.05)
print(rg.Pattern(synth.DATE_CAPTURE_PATTERN).find_reader_offsets(s, rg.StreamConfig(0, 0))[0])
print(rg.Pattern(r'(?P<y>\d{4})-(?P<m>\d+)').find_reader_offsets(s, rg.StreamConfig(0, 0))[0])
print(p.find_reader_offsets(synth.make_buffer('url', 300000), rg.StreamConfig(0, 0))[0])
print(len(p2.replace_all(bytes(synth.make_buffer('log', 100000)), '[\$user]')))
" > $out/racecheck.log 2>&1; echo "racecheck rc=$?" >> $out/racecheck.log
tail -5 $out/racecheck.log
