#!/bin/bash
# 1-GPU call, kernels v22 (phase-D ranking with 256 buckets, per-bin list spill, 8 bins per CTA from R = 8): GPU tests, bench lines
# of every configuration, sanitizers, launch lists, full ncu captures, phase probes
O=gpurun_out/r02v; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
for c in blockout general buffered10 cube1 general24; do
  timeout 400 python bench.py --steps 20 --warmup 5 --config $c --cpu-seconds 9 > $O/bench_$c.json 2> $O/bench_$c.err
done
timeout 400 python bench.py --steps 50 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/ref_blockout.json 2> $O/ref_blockout.err
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -x -q \
   -k "episode_matches_reference_golden or buffered or all_possible or hull_fixture or item_generator or reloaded or 24_rotations" > $O/memcheck.log 2>&1; echo "memcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $O/memcheck.log | tail -2 | tr '\n' ' ')" | tee -a $O/summary.txt
timeout 500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "episode_blockout or episode_truncate or episode_irregular or buffered_episode" > $O/racecheck.log 2>&1; echo "racecheck rc=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' $O/racecheck.log | tail -2 | tr '\n' ' ')" | tee -a $O/summary.txt
for w in blockout irregular8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches_$w.csv python tools/kbench.py --workloads $w --steps 20 --burn 140 > /dev/null 2>> $O/err.txt
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:irbpp --launch-skip 300 -c 2 -o $O/prof_v22_$w -f \
     python tools/kbench.py --workloads $w --steps 10 --burn 160 > /dev/null 2>> $O/err.txt
done
IRBPP_PROBE_CONFIG=blockout IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_fine.so timeout 200 python tools/phase_probe.py > $O/phase_blockout.json 2>> $O/err.txt
IRBPP_PROBE_CONFIG=general IRBPP_LIB=$PWD/irbpp_b200/lib/libirbpp_fine.so timeout 200 python tools/phase_probe.py > $O/phase_general.json 2>> $O/err.txt
timeout 200 python tools/kbench.py --workloads blockout,irregular8,irregular24,buffered10,cube --steps 40 > $O/kbench.jsonl 2>> $O/err.txt
cat $O/summary.txt; cat $O/kbench.jsonl; for c in default blockout general buffered10 cube1 general24; do grep '^{' $O/bench_$c.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$c value %.4e ms/step %.4f e2e %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac']), d.get('step_ms'), 'cpu', d.get('cpu_baseline',{}).get('value'))
"; done; tail -3 $O/err.txt
