#!/bin/bash
# GPU call 1 of round 2: parity of the default build, then every prepared variant timed on BlockOut and irregular R=8.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_default.log 2>&1; echo "pytest default rc=$?" | tee -a $O/summary.txt
L=irbpp_b200/lib
for v in "" _coop9 _coop13 _coop17 _split _split10 _both13 _pf; do
  export IRBPP_LIB=$PWD/$L/libirbpp$v.so
  if [ -n "$v" ]; then
    timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "episode or random or hull" > $O/pytest$v.log 2>&1
    echo "variant $v parity rc=$? $(tail -1 $O/pytest$v.log)" | tee -a $O/summary.txt
  fi
  timeout 300 python tools/kbench.py --workloads blockout,irregular8 --e2e >> $O/kbench.jsonl 2>> $O/kbench.err
done
unset IRBPP_LIB
for hr in kernel memcpy; do
  IRBPP_HOST_RESULTS=$hr timeout 300 python tools/kbench.py --workloads blockout --e2e >> $O/kbench_e2e.jsonl 2>> $O/kbench.err
  IRBPP_HOST_RESULTS=$hr IRBPP_HOST_ACTIONS=memcpy timeout 300 python tools/kbench.py --workloads blockout --e2e >> $O/kbench_e2e.jsonl 2>> $O/kbench.err
done
IRBPP_HOST_ACTIONS=memcpy timeout 300 python tools/kbench.py --workloads blockout --e2e >> $O/kbench_e2e.jsonl 2>> $O/kbench.err
# per-kernel durations (launch lists) of default / split / coop13 on both workloads
for v in "" _split _coop13; do
  for w in blockout irregular8; do
  IRBPP_LIB=$PWD/$L/libirbpp$v.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches${v}_$w.csv python tools/kbench.py --workloads $w --steps 20 --burn 140 > /dev/null 2>> $O/kbench.err
  done
done
cat $O/kbench.jsonl $O/kbench_e2e.jsonl
