#!/bin/bash
# GPU call 3 of round 2: v17 kernels (4-wide serial approxPolyDP scan, run-length border follower) + the round-2 additions
O=gpurun_out/r02c; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
timeout 600 python tools/kbench.py --workloads blockout,irregular8,cube,buffered10 --e2e > $O/kbench.jsonl 2> $O/kbench.err
for w in blockout irregular8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches_$w.csv python tools/kbench.py --workloads $w --steps 20 --burn 140 > /dev/null 2>> $O/kbench.err
done
for c in blockout general; do
  IRBPP_PROBE_CONFIG=$c IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_fine.so timeout 300 python tools/phase_probe.py > $O/phase_$c.json 2>> $O/kbench.err
done
timeout 300 python tools/actor_loop.py --iters 40 > $O/actor_loop.json 2>> $O/kbench.err
cat $O/kbench.jsonl; cat $O/phase_*.json; cat $O/actor_loop.json
