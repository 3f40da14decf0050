#!/bin/bash
# GPU call 8 of round 2: v20 (phase-A prefetch) -- tests, all workloads, launch lists, full ncu captures for profiles/
O=gpurun_out/r02h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
timeout 600 python tools/kbench.py --workloads blockout,irregular8,cube,buffered10 --e2e > $O/kbench.jsonl 2> $O/err.txt
for w in blockout irregular8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches_$w.csv python tools/kbench.py --workloads $w --steps 20 --burn 140 > /dev/null 2>> $O/err.txt
done
IRBPP_PROBE_CONFIG=blockout IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_fine.so timeout 300 python tools/phase_probe.py > $O/phase_blockout.json 2>> $O/err.txt
IRBPP_PROBE_CONFIG=general IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_fine.so timeout 300 python tools/phase_probe.py > $O/phase_general.json 2>> $O/err.txt
timeout 300 python tools/actor_loop.py --iters 40 > $O/actor_loop.json 2>> $O/err.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:irbpp --launch-skip 300 -c 2 -o $O/prof_v20_blockout -f \
   python tools/kbench.py --workloads blockout --steps 10 --burn 160 > /dev/null 2>> $O/err.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:irbpp --launch-skip 300 -c 2 -o $O/prof_v20_irregular8 -f \
   python tools/kbench.py --workloads irregular8 --steps 10 --burn 160 > /dev/null 2>> $O/err.txt
cat $O/kbench.jsonl; cat $O/phase_blockout.json; cat $O/actor_loop.json
