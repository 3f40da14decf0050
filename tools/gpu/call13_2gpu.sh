#!/bin/bash
# 2-GPU call: CUDA-graph host step, compact rollout gather, bench runway / NVML sampler
O=gpurun_out/r02m; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $O/pytest.log)" | tee $O/summary.txt
timeout 300 python tools/kbench.py --workloads blockout,irregular8 --e2e > $O/kbench_graph.jsonl 2> $O/err.txt
IRBPP_GRAPH=0 timeout 300 python tools/kbench.py --workloads blockout --e2e > $O/kbench_nograph.jsonl 2>> $O/err.txt
timeout 300 python tools/e2e_probe.py 200 > $O/e2e_probe.txt 2>> $O/err.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_compact.json 2> $O/bench_n2_compact.err
IRBPP_GATHER=nccl timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
cat $O/kbench_graph.jsonl $O/kbench_nograph.jsonl; cat $O/e2e_probe.txt
for f in bench_n1 bench_n2_compact bench_n2_nccl; do echo $f; grep '^{' $O/$f.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  value %.4e ms/step %.4f kern %.4f e2e %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['ms_per_step']), d.get('step_ms'), d.get('clocks'), {k:(round(d[k],3) if isinstance(d[k],float) else d[k]) for k in ('gather_ms','gather_exposed_ms','gather_hidden_ms','gather_alone_ms','gather_kind') if d.get(k) is not None})
"; tail -1 $O/$f.err | cut -c1-200; done
