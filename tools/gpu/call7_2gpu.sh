#!/bin/bash
# 2-GPU call: bench.py N = 2 (pipelined rollout gather over NCCL), the reference arm under torchrun, the general config at N = 2, gpu tests needing 2 devices
O=gpurun_out/r02g; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --config abc32k > $O/bench_abc_n2.json 2> $O/bench_abc_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 python -m pytest tests/test_gpu_parity2.py -m gpu -x -q -k "two_devices" > $O/pytest_2dev.log 2>&1; echo "2dev rc=$? $(tail -1 $O/pytest_2dev.log)" | tee $O/summary.txt
for f in bench_n1 bench_n2 bench_abc_n2 bench_ref_n2; do echo $f; cut -c1-900 $O/$f.json; tail -2 $O/$f.err; done
