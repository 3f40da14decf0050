#!/bin/bash
# 1-GPU call: 7 (and 6) bins per CTA for the wide shape of the candidates kernel: 4096 / 7 = 586 CTAs fit ONE wave at 4 CTAs/SM
O=gpurun_out/r02y; mkdir -p $O
L=$PWD/irbpp_b200/lib
for v in default e7 e6 default e7; do
  f=$L/libirbpp_$v.so; [ $v = default ] && f=$L/libirbpp.so
  IRBPP_LIB=$f timeout 150 python tools/kbench.py --workloads irregular8,irregular24 --steps 50 --burn 120 >> $O/sweep.jsonl 2>> $O/err.txt
done
IRBPP_LIB=$L/libirbpp_e7.so timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_e7.log 2>&1; echo "pytest e7 rc=$? $(tail -1 $O/pytest_e7.log)" | tee $O/summary.txt
IRBPP_LIB=$L/libirbpp_e7.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:irbpp --launch-skip 300 -c 2 -o $O/prof_e7_irregular8 -f \
     python tools/kbench.py --workloads irregular8 --steps 10 --burn 160 > /dev/null 2>> $O/err.txt
IRBPP_LIB=$L/libirbpp_e7.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:irbpp --launch-skip 250 -c 60 --csv \
     --log-file $O/launches_irregular8.csv python tools/kbench.py --workloads irregular8 --steps 20 --burn 140 > /dev/null 2>> $O/err.txt
IRBPP_LIB=$L/libirbpp_e7.so timeout 300 python bench.py --steps 20 --warmup 5 --config general --no-cpu-baseline > $O/bench_general_e7.json 2> $O/bench_general_e7.err
cat $O/summary.txt; cat $O/sweep.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%-22s %-11s %.4f min %.4f p90 %.4f'%(d['lib'],d['workload'],d['ms_per_step'],d['ms_min'],d['ms_p90']))"
grep '^{' $O/bench_general_e7.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('general e7 value %.4e ms/step %.4f e2e %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['roofline']['frac']))"
tail -3 $O/err.txt
