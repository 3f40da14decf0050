#!/bin/bash
# 8-GPU call: how many NCCL channels (SMs) should the pipelined rollout all-gather take?  N = 8 sweep, then N = 4 with the same settings
O=gpurun_out/r02l; mkdir -p $O
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/n1.json 2> $O/n1.err
for ch in 0 16 8 4 2; do
  IRBPP_NCCL_CHANNELS=$ch timeout 300 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > $O/n8_ch$ch.json 2> $O/n8_ch$ch.err
done
for f in n1 n8_ch0 n8_ch16 n8_ch8 n8_ch4 n8_ch2; do echo $f; grep '^{' $O/$f.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  value %.4e ms/step %.4f kern %.4f e2e %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['ms_per_step']), d.get('step_ms'), {k:(round(d[k],3) if isinstance(d[k],float) else d[k]) for k in ('gather_ms','gather_exposed_ms','gather_hidden_ms','gather_alone_ms','nccl_channels') if d.get(k) is not None})
"; tail -1 $O/$f.err | cut -c1-200; done
