#!/bin/bash
out=gpurun_out/${1:-misc}; mkdir -p $out
cap() { timeout 120 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $out/full_$1 python tools/prof_misc.py $4 $5 > $out/ncu_$1.log 2>&1; ls -la $out/full_$1.ncu-rep 2>/dev/null | awk '{print $5, $9}'; }
cap linear findall_scan_linear 2 linear 256
cap generic "findall_scan_kernel" 2 generic 64
cap replace1 "replace_batch_kernel<\(int\)1" 2 replace 32
cap table find_reader_table_kernel 2 table 128
