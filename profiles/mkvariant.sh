#!/bin/bash
# usage: profiles/mkvariant.sh NAME "N list" "extra nvcc defines"   -> build/variants/libsympa_NAME.so
set -e
cd /root/repo
NAME=$1; NS=$2; DEFS=$3
OBJS="build/obj/sympa_b200.o"
for n in 1 2 3 4 5 6 7 8 9 10; do
  if [[ " $NS " == *" $n "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DSYMPA_TU_N=$n $DEFS -Xptxas -v -c sympa_b200/csrc/pair_kernels_n.cu -o build/variants/${NAME}_$n.o 2> build/variants/${NAME}_$n.log &
    OBJS="$OBJS build/variants/${NAME}_$n.o"
  else
    OBJS="$OBJS build/obj/pair_kernels_$n.o"
  fi
done
wait
nvcc -shared -o build/variants/libsympa_$NAME.so $OBJS -gencode arch=compute_100a,code=sm_100a
echo build/variants/libsympa_$NAME.so
