#!/bin/bash
# 1-GPU call: bench lines of every BASELINE configuration (own arm + CPU arm), GPU tests, compute-sanitizer on the v20 kernels
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
for c in blockout general buffered10 cube1 general24; do
  timeout 600 python bench.py --steps 20 --warmup 5 --config $c > $O/bench_$c.json 2> $O/bench_$c.err
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/ref_blockout.json 2> $O/ref_blockout.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --config general > $O/ref_general.json 2> $O/ref_general.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --config buffered10 > $O/ref_buffered10.json 2> $O/ref_buffered10.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --config cube1 > $O/ref_cube1.json 2> $O/ref_cube1.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -x -q \
   -k "episode_matches_reference_golden or buffered or all_possible or hull_fixture or item_generator or reloaded" > $O/memcheck.log 2>&1; echo "memcheck rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $O/memcheck.log | tail -2 | tr '\n' ' ')" | tee -a $O/summary.txt
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
   -k "episode_blockout or episode_truncate or buffered_episode" > $O/racecheck.log 2>&1; echo "racecheck rc=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' $O/racecheck.log | tail -2 | tr '\n' ' ')" | tee -a $O/summary.txt
cat $O/summary.txt; for c in blockout general buffered10 cube1 general24; do grep '^{' $O/bench_$c.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$c value %.4e ms/step %.4f e2e %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac']), d.get('step_ms'), 'cpu', d.get('cpu_baseline',{}).get('value'))
"; done
