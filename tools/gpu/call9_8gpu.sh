#!/bin/bash
# 8-GPU call: the 1 -> 8 curve of bench.py (pipelined rollout gather), abc32k (BASELINE.json configs[4]) at 8 GPUs, reference arm
O=gpurun_out/r02i; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
run() { n=$1; shift; if [ $n -eq 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@"; fi; }
for n in 1 2 4 8; do
  timeout 600 bash -c "$(declare -f run); run $n --steps 20 --warmup 5 --no-cpu-baseline" > $O/bench_n$n.json 2> $O/bench_n$n.err
done
timeout 600 bash -c "$(declare -f run); run 8 --steps 20 --warmup 5 --config abc32k" > $O/bench_abc_n8.json 2> $O/bench_abc_n8.err
timeout 600 bash -c "$(declare -f run); run 8 --impl reference --steps 20 --warmup 5" > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err
timeout 600 bash -c "$(declare -f run); run 8 --impl reference --steps 20 --warmup 5 --config abc32k" > $O/bench_ref_abc_n8.json 2> $O/bench_ref_abc_n8.err
for f in bench_n1 bench_n2 bench_n4 bench_n8 bench_abc_n8 bench_ref_n8 bench_ref_abc_n8; do echo $f; grep '^{' $O/$f.json | cut -c1-300; tail -1 $O/$f.err; done
