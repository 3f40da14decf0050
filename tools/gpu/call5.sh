#!/bin/bash
# GPU call 5 of round 2: v19 (3 tasks per lane per batch, phase-D lists in global scratch for R > 4, run-time bins per CTA)
O=gpurun_out/r02e; mkdir -p $O
timeout 600 python tools/kbench.py --workloads blockout,irregular8,cube,irregular24,buffered10 --e2e > $O/kbench.jsonl 2> $O/kbench.err
for b in 1 2 4; do IRBPP_BINS_PER_CTA=$b timeout 300 python tools/kbench.py --workloads blockout,irregular8 --steps 40 >> $O/kbench_bins.jsonl 2>> $O/kbench.err; echo "bins_per_cta=$b" >> $O/kbench_bins.jsonl; done
for w in blockout irregular8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches_$w.csv python tools/kbench.py --workloads $w --steps 20 --burn 140 > /dev/null 2>> $O/kbench.err
done
IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_trace.so timeout 300 python tools/cta_trace.py blockout > $O/trace_blockout.json 2>> $O/kbench.err
IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_trace.so timeout 300 python tools/cta_trace.py general > $O/trace_general.json 2>> $O/kbench.err
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
timeout 300 python tools/e2e_probe.py 200 > $O/e2e_probe.txt 2>> $O/kbench.err
timeout 300 python tools/actor_loop.py --iters 40 > $O/actor_loop.json 2>> $O/kbench.err
cat $O/kbench.jsonl $O/kbench_bins.jsonl; cat $O/e2e_probe.txt; cat $O/actor_loop.json
