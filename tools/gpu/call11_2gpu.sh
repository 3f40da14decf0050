#!/bin/bash
O=gpurun_out/r02k; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 300 $TR tools/p2p_probe.py > $O/p2p.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
cat $O/p2p.txt | grep -v "^\*\|OMP_NUM"; for f in bench_n1 bench_n2_nccl; do echo $f; grep '^{' $O/$f.json | cut -c1-200; tail -1 $O/$f.err; done
