#!/bin/bash
# what the driver runs at round end, on one B200: GPU tests, smoke, the default bench line and the reference arm
cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
( time python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err ) 2>&1 | grep real
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench_n1.err ) 2>&1 | grep real
python -c "
import json
d=json.load(open('gpurun_out/final_bench_n1.json')); r=json.load(open('gpurun_out/final_bench_ref.json'))
print('value %.4e ms/step %.2f e2e %.4e frac %.4f n10 %.4e launches %d clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['n10']['value'], d['gpu_launches'], d['clocks']))
print('reference arm %.4e pairs/s' % r['value'], r['cpu_baseline']['kind'], 'cpu_baseline', d['cpu_baseline']['value'])
print('cfg4', d['also']['cfg4_1Mnodes_upper_fmin_n10_train_step_ms_dp1'], d['also']['cfg4_1Mnodes_spd_n10_train_step_ms_dp1'])"
