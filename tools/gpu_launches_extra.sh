#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'find_reader|exclusive' -c 60 --csv --log-file $out/launches_reader.csv \
   python tools/bench_extra.py --max-patterns 0 --stream-gib 0.25 > $out/ncu_reader.log 2>&1
tail -2 $out/ncu_reader.log | cut -c1-300
