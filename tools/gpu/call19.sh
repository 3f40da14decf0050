#!/bin/bash
# 1-GPU call: CTA-shape sweep of both kernels (bins per CTA, warps per CTA, task slots per lane, scan residency)
O=gpurun_out/r02s; mkdir -p $O
L=$PWD/irbpp_b200/lib
: > $O/sweep.jsonl
for v in default e2w4 e1w4 e2w2 e8w8 e4w8 tpl1 tpl2 scan7 scan9 e2w4t1 e2w4t2 e1w4t1; do
  f=$L/libirbpp_$v.so; [ $v = default ] && f=$L/libirbpp.so
  IRBPP_LIB=$f timeout 120 python tools/kbench.py --workloads blockout,irregular8 --steps 40 --burn 120 >> $O/sweep.jsonl 2>> $O/err.txt
done
cat $O/sweep.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-22s %-11s %.4f min %.4f p90 %.4f'%(d['lib'],d['workload'],d['ms_per_step'],d['ms_min'],d['ms_p90']))"
tail -3 $O/err.txt
