#!/bin/bash
# 2-GPU call: copy-engine gather over symmetric memory (probe + bench), send/recv pairs
O=gpurun_out/r02r; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 200 $TR tools/p2p_probe.py > $O/p2p_probe.txt 2> $O/p2p_probe.err
for g in symm nccl; do
  IRBPP_GATHER=$g timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n2_$g.json 2> $O/bench_n2_$g.err
done
NCCL_P2P_USE_CUDA_MEMCPY=1 IRBPP_GATHER=sendrecv timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_n2_sendrecv_ce.json 2> $O/bench_n2_sendrecv_ce.err
cat $O/p2p_probe.txt; tail -3 $O/p2p_probe.err
for f in bench_n2_symm bench_n2_nccl bench_n2_sendrecv_ce; do echo $f; grep '^{' $O/$f.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  value %.4e ms/step %.4f kern %.4f e2e %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['ms_per_step']), d.get('step_ms'), {k:(round(d[k],3) if isinstance(d[k],float) else d[k]) for k in ('gather_ms','gather_exposed_ms','gather_hidden_ms','gather_alone_ms','gather_kind') if d.get(k) is not None})
"; tail -2 $O/$f.err | cut -c1-300; done
