#!/bin/bash
# 1-GPU call: per-bin hand-over flags between the scan and the candidates kernel (instead of the grid-wide dependency wait)
O=gpurun_out/r02w; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
for rep in 1 2; do
  timeout 200 python tools/kbench.py --workloads blockout,irregular8,buffered10 --steps 60 >> $O/kbench_flags.jsonl 2>> $O/err.txt
  IRBPP_HANDOVER=grid timeout 200 python tools/kbench.py --workloads blockout,irregular8,buffered10 --steps 60 >> $O/kbench_grid.jsonl 2>> $O/err.txt
done
timeout 400 python bench.py --steps 50 --warmup 5 --cpu-seconds 6 > $O/bench_default.json 2> $O/bench_default.err
IRBPP_HANDOVER=grid timeout 400 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_grid.json 2> $O/bench_grid.err
timeout 400 python bench.py --steps 20 --warmup 5 --config general --no-cpu-baseline > $O/bench_general.json 2> $O/bench_general.err
for w in blockout irregular8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches_$w.csv python tools/kbench.py --workloads $w --steps 20 --burn 140 > /dev/null 2>> $O/err.txt
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:irbpp --launch-skip 300 -c 2 -o $O/prof_v23_$w -f \
     python tools/kbench.py --workloads $w --steps 10 --burn 160 > /dev/null 2>> $O/err.txt
done
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "episode_matches_reference_golden or buffered" > $O/memcheck.log 2>&1; echo "memcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $O/memcheck.log | tail -2 | tr '\n' ' ')" | tee -a $O/summary.txt
cat $O/summary.txt; echo flags; cat $O/kbench_flags.jsonl; echo grid; cat $O/kbench_grid.jsonl
for c in default grid general; do grep '^{' $O/bench_$c.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$c value %.4e ms/step %.4f e2e %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac']), d.get('step_ms'))
"; done; tail -3 $O/err.txt
