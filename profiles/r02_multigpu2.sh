#!/bin/bash
# 8-GPU follow-up: concurrent H2D bandwidth, then the bench line (two calls per rank, accumulated + overlapped)
cd $GRAFT_REPO_ROOT
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/r02_h2d.py 2>/dev/null | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print('value %.4e ms/step %.3f e2e %.4e (%.3f ms) chunk %d frac %.4f kern %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['chunk_pairs'], d['roofline']['frac'], d['roofline']['kernel_ms']))
print('parity', d.get('parity_check',{}).get('ok'), d.get('parity_check',{}).get('max_rel'))
print('n10 %.4e e2e %.4e' % (d['n10']['value'], d['n10']['e2e']['value']))
print('launches', d['gpu_launches'], d['gpu_launches_how'][:60])
a=d.get('also',{})
for k in a:
    if 'cfg' in k and 'ms' in k or 'config3' in k and 'sec' in k: print(' ', k, a[k])
"
grep -v "UserWarning\|return func\|^$\|\*\*\*\*\|OMP_NUM\|torch.det" gpurun_out/r02_bench_n$N.err | tail -4
