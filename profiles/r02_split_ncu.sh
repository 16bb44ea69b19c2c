cd $GRAFT_REPO_ROOT
Q="--steps 1 --warmup 3 --no-cpu-baseline --no-extras --no-n10"
SYMPA_SPLIT_PATH=1 ncu --set full --clock-control none --import-source on -k regex:coop_spectrum -s 6 -c 1 -o gpurun_out/r02_spectrum_n10 -f python bench.py --n 10 --metric fmin $Q --pairs 524288 > gpurun_out/r02_split_ncu2.log 2>&1
tail -2 gpurun_out/r02_split_ncu2.log | cut -c1-200
