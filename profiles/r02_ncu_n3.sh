cd $GRAFT_REPO_ROOT
Q="--steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-n10"
ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 2 -c 1 -o gpurun_out/r02_pair_n3 -f python bench.py --n 3 $Q --pairs 2097152 > gpurun_out/r02_ncu_n3.log 2>&1
tail -1 gpurun_out/r02_ncu_n3.log | cut -c1-100
