#!/bin/bash
# 2-GPU call: the copy-engine peer gather against the NCCL all-gather, N = 1 rehearsal check, byte-offset scan lists
O=gpurun_out/r02j; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_peer.json 2> $O/bench_n2_peer.err
IRBPP_GATHER=nccl timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --config abc32k > $O/bench_abc_n2_peer.json 2> $O/bench_abc_n2.err
timeout 300 python tools/kbench.py --workloads blockout,irregular8 > $O/kbench.jsonl 2> $O/err.txt
for f in bench_n1 bench_n2_peer bench_n2_nccl bench_abc_n2_peer; do echo $f; grep '^{' $O/$f.json | cut -c1-200; tail -2 $O/$f.err; done; cat $O/kbench.jsonl
