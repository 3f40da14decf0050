#!/bin/bash
# 1-GPU call: 24 rotations after the slot-cap change (default, 1 task per lane), overflow-redo counters
O=gpurun_out/r02o; mkdir -p $O
L=$PWD/irbpp_b200/lib
timeout 300 python tools/kbench.py --workloads blockout,irregular8,irregular24 --steps 40 > $O/kbench.jsonl 2> $O/err.txt
IRBPP_LIB=$L/libirbpp_tpl1.so timeout 300 python tools/kbench.py --workloads irregular8,irregular24 --steps 40 > $O/kbench_tpl1.jsonl 2>> $O/err.txt
for c in blockout general general24; do IRBPP_PROBE_CONFIG=$c timeout 300 python tools/phase_probe.py > $O/phase_$c.json 2>> $O/err.txt; done
cat $O/kbench.jsonl $O/kbench_tpl1.jsonl; grep -h "overflow\|config\|tasks\|images" $O/phase_*.json
