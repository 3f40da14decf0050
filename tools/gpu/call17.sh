#!/bin/bash
# 1-GPU call: phase-D ranking with 256 adaptive height buckets (R = 8 and R = 24)
O=gpurun_out/r02q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
timeout 300 python tools/kbench.py --workloads blockout,irregular8,irregular24,buffered10 --steps 40 > $O/kbench.jsonl 2> $O/err.txt
IRBPP_PROBE_CONFIG=general timeout 300 python tools/phase_probe.py > $O/phase_general.json 2>> $O/err.txt
IRBPP_PROBE_CONFIG=general24 timeout 300 python tools/phase_probe.py > $O/phase_general24.json 2>> $O/err.txt
cat $O/kbench.jsonl; cat $O/phase_general.json $O/phase_general24.json
