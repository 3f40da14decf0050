#!/bin/bash
# 1-GPU call: phase-D lists of R = 8 in the 4 KB warp scratch with a global spill (7 CTAs/SM instead of 4), also with 8 bins per CTA
O=gpurun_out/r02u; mkdir -p $O
L=$PWD/irbpp_b200/lib
: > $O/sweep.jsonl
for v in default l4 l4e8 e8w8; do
  f=$L/libirbpp_$v.so; [ $v = default ] && f=$L/libirbpp.so
  IRBPP_LIB=$f timeout 150 python tools/kbench.py --workloads blockout,irregular8,irregular24 --steps 40 --burn 120 >> $O/sweep.jsonl 2>> $O/err.txt
done
IRBPP_LIB=$L/libirbpp_l4.so timeout 400 python -m pytest tests -m gpu -x -q -k "rot24 or 24_rotations or truncate or irregular or general or 32768 or episode" > $O/pytest_l4.log 2>&1; echo "pytest l4 rc=$? $(tail -1 $O/pytest_l4.log)" | tee $O/summary.txt
cat $O/sweep.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-22s %-11s %.4f min %.4f p90 %.4f'%(d['lib'],d['workload'],d['ms_per_step'],d['ms_min'],d['ms_p90']))"
tail -3 $O/err.txt
