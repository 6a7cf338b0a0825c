#!/bin/bash
# Run HERE after tools/capture_profiles.sh came back from gpurun: copies the on-box summaries from gpurun_out/ into
# the tracked profiles/ directory and adds the SASS opcode histogram of the shipped library.
R=gpurun_out; P=profiles
cp $R/r02_*_ncu.csv $R/r02_*_lines.txt $R/r02_bench_*.json $R/r02_launches.csv $R/r02_bpr_large.json $R/r02_eval_large.json $R/r02_bpr_c2_means.txt $P/ 2>/dev/null
[ -f $R/traffic.json ] && cp $R/traffic.json $P/traffic.json
cuobjdump -sass qmf_b200/libqmf_b200.so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn > $P/r02_sass_opcodes.txt
cuobjdump -sass qmf_b200/libqmf_b200.so | grep -oE "(DMMA|UBLKCP|LDGSTS|SYNCS|UTMALDG|REDG|RED|ATOMG)[A-Z0-9_.]*" | sort | uniq -c | sort -rn > $P/r02_sass_key_opcodes.txt
ls -la $P | tail -40
