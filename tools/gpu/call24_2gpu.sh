#!/bin/bash
# 2-GPU sanity run of the final state: bench default (NCCL all-gather beside the v23 kernels) + the reference arm under torchrun
O=gpurun_out/r02x; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --config general > $O/bench_general_n2.json 2> $O/bench_general_n2.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > $O/ref_n2.json 2> $O/ref_n2.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1
for f in bench_n2 bench_general_n2; do echo $f; grep '^{' $O/$f.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  value %.4e strict %.4e ms/step %.4f kern %.4f e2e %.4f traffic %s'%(d['value'],d.get('value_charging_flush_gaps',0),d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['ms_per_step'],d['roofline']['traffic']), d.get('step_ms'), {k:(round(d[k],3) if isinstance(d[k],float) else d[k]) for k in ('gather_ms','gather_exposed_ms','gather_hidden_ms','gather_beside_flush_ms','gather_tail_ms','gather_alone_ms','gather_kind') if d.get(k) is not None})
"; tail -2 $O/$f.err | cut -c1-300; done
grep '^{' $O/ref_n2.json | cut -c1-300; tail -2 $O/smoke.txt
