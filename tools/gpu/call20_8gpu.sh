#!/bin/bash
# 8-GPU call: rollout gather by copy engines over symmetric memory vs NCCL all-gather vs packed all-gather; 1 / 4 / 8 GPUs; abc32k
O=gpurun_out/r02t; mkdir -p $O
run() { n=$1; shift; if [ $n -eq 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $n "$@"; fi; }
go() { name=$1; n=$2; shift 2; timeout 170 bash -c "$(declare -f run); run $n --steps 20 --warmup 5 --no-cpu-baseline $*" > $O/$name.json 2> $O/$name.err; }
IRBPP_GATHER=symm go bench_n8_symm 8
IRBPP_GATHER=nccl go bench_n8_nccl 8
IRBPP_GATHER=compact go bench_n8_compact 8
go bench_n4 4
go bench_n1 1
go bench_abc_n8 8 --config abc32k
for f in bench_n8_symm bench_n8_nccl bench_n8_compact bench_n4 bench_n1 bench_abc_n8; do echo $f; grep '^{' $O/$f.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  value %.4e strict %.4e ms/step %.4f kern %.4f e2e %.4f'%(d['value'],d.get('value_charging_flush_gaps',0),d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['ms_per_step']), d.get('step_ms'), {k:(round(d[k],3) if isinstance(d[k],float) else d[k]) for k in ('gather_ms','gather_exposed_ms','gather_hidden_ms','gather_beside_flush_ms','gather_tail_ms','gather_alone_ms','gather_kind') if d.get(k) is not None})
"; tail -2 $O/$f.err | cut -c1-300; done
