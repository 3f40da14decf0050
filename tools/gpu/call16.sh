#!/bin/bash
# 1-GPU call: R = 24 with phase-D lists in global memory (7 CTAs/SM instead of 2)
O=gpurun_out/r02p; mkdir -p $O
L=$PWD/irbpp_b200/lib
timeout 600 python -m pytest tests -m gpu -x -q -k "rot24 or 24_rotations or two_handles or truncate or irregular" > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
timeout 300 python tools/kbench.py --workloads blockout,irregular8,irregular24 --steps 40 > $O/kbench.jsonl 2> $O/err.txt
IRBPP_LIB=$L/libirbpp_tpl1.so timeout 300 python tools/kbench.py --workloads irregular24 --steps 40 > $O/kbench_tpl1.jsonl 2>> $O/err.txt
IRBPP_PROBE_CONFIG=general24 timeout 300 python tools/phase_probe.py > $O/phase_general24.json 2>> $O/err.txt
cat $O/kbench.jsonl $O/kbench_tpl1.jsonl; cat $O/phase_general24.json
