#!/bin/bash
# GPU call 2 of round 2: v16 kernels (packed approxPolyDP, bulk-copy staging, in-place write-back, dense general scan)
O=gpurun_out/r02b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
timeout 600 python tools/kbench.py --workloads blockout,irregular8,cube,irregular24,buffered10 --e2e > $O/kbench.jsonl 2> $O/kbench.err
for w in blockout irregular8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches_$w.csv python tools/kbench.py --workloads $w --steps 20 --burn 140 > /dev/null 2>> $O/kbench.err
done
for c in blockout general; do
  IRBPP_PROBE_CONFIG=$c IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_fine.so timeout 300 python tools/phase_probe.py > $O/phase_$c.json 2>> $O/kbench.err
done
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_blockout.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2>> $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:irbpp --launch-skip 300 -c 2 -o $O/prof_v16_blockout -f \
   python tools/kbench.py --workloads blockout --steps 10 --burn 160 > /dev/null 2>> $O/kbench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:irbpp --launch-skip 300 -c 2 -o $O/prof_v16_irregular8 -f \
   python tools/kbench.py --workloads irregular8 --steps 10 --burn 160 > /dev/null 2>> $O/kbench.err
cat $O/kbench.jsonl; cat $O/bench_blockout.json | cut -c1-600; cat $O/bench_ref.json | cut -c1-300
