#!/bin/bash
# GPU call 4 of round 2: v18 (follower back to the r01 form, PPL-chunked dense scan, host actions by memcpy) + per-CTA timelines
O=gpurun_out/r02d; mkdir -p $O
timeout 600 python tools/kbench.py --workloads blockout,irregular8 --e2e > $O/kbench.jsonl 2> $O/kbench.err
for w in blockout irregular8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches_$w.csv python tools/kbench.py --workloads $w --steps 20 --burn 140 > /dev/null 2>> $O/kbench.err
done
IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_trace.so timeout 300 python tools/cta_trace.py blockout > $O/trace_blockout.json 2>> $O/kbench.err
IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_trace.so timeout 300 python tools/cta_trace.py general > $O/trace_general.json 2>> $O/kbench.err
IRBPP_PROBE_CONFIG=blockout IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_fine.so timeout 300 python tools/phase_probe.py > $O/phase_blockout.json 2>> $O/kbench.err
timeout 600 python -m pytest tests -m gpu -x -q -k "episode or random or hull or buffered" > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
cat $O/kbench.jsonl; cat $O/trace_blockout.json | head -80
