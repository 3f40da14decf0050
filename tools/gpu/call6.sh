#!/bin/bash
# GPU call 6 of round 2: tasks-per-lane x list placement x bins-per-CTA sweep (general + BlockOut), host path after the double-buffered result block
O=gpurun_out/r02f; mkdir -p $O
L=$PWD/irbpp_b200/lib
for t in "" _tpl2 _tpl1; do
  for lists in global smem; do
    for b in 2 4; do
      IRBPP_LIB=$L/libirbpp$t.so IRBPP_LISTS=$lists IRBPP_BINS_PER_CTA=$b timeout 200 python tools/kbench.py --workloads irregular8 --steps 60 >> $O/sweep.jsonl 2>> $O/err.txt
      echo "{\"tpl\": \"$t\", \"lists\": \"$lists\", \"bins\": $b}" >> $O/sweep.jsonl
    done
  done
  IRBPP_LIB=$L/libirbpp$t.so timeout 200 python tools/kbench.py --workloads blockout,cube --steps 60 >> $O/sweep.jsonl 2>> $O/err.txt
  echo "{\"tpl\": \"$t\"}" >> $O/sweep.jsonl
done
timeout 300 python tools/kbench.py --workloads blockout,irregular8 --e2e > $O/kbench.jsonl 2>> $O/err.txt
timeout 300 python tools/e2e_probe.py 200 > $O/e2e_probe.txt 2>> $O/err.txt
timeout 300 python tools/actor_loop.py --iters 40 > $O/actor_loop.json 2>> $O/err.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "contract or episode or reset_specific or c_abi or actor" > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
cat $O/sweep.jsonl | cut -c1-200; cat $O/kbench.jsonl; cat $O/e2e_probe.txt; cat $O/actor_loop.json
