#!/bin/bash
# multi-GPU check: gpurun --gpus N -- 'bash profiles/r02_multigpu.sh N'
cd $GRAFT_REPO_ROOT
N=${1:-2}
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r02_topo_n$N.txt; for f in /sys/devices/system/node/node*/cpulist; do echo "$f $(cat $f)"; done >> gpurun_out/r02_topo_n$N.txt; nproc >> gpurun_out/r02_topo_n$N.txt
python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
grep -v "UserWarning\|return func\|^$\|\*\*\*\*\|OMP_NUM" gpurun_out/r02_bench_n$N.err | tail -8
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print('value %.4e ms/step %.3f e2e %.4e (%.3f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
print('parity', d.get('parity_check',{}).get('ok'), d.get('parity_check',{}).get('max_rel'))
print('n10 %.4e e2e %.4e' % (d['n10']['value'], d['n10']['e2e']['value']))
print('launches', d['gpu_launches'], d['gpu_launches_how'][:90])
for k,v in d.get('also',{}).items(): print(' ', k, v)
"
# A/B: the same short run without the NUMA binding of the ranks
for v in 1; do
  SYMPA_BENCH_NO_NUMA=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus $N --steps 10 --warmup 3 --no-extras --no-n10 > gpurun_out/r02_numa${v}_n$N.json 2>> gpurun_out/r02_bench_n$N.err
  python -c "
import json
d=json.load(open('gpurun_out/r02_numa${v}_n$N.json'))
print('no_numa=$v value %.4e ms/step %.3f e2e %.4e (%.3f ms) enqueue %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('host_enqueue_ms_per_step', -1)), d['config']['host_numa'][:60])"
done
