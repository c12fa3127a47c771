#!/bin/bash
# Role-ablation builds of the tcgen05 EdgeConv backward kernel (timing experiments only: every variant but 0 computes garbage; the
# forward kernel of round 2, which had the same switches, was replaced in round 3 — its timings are in profiles/r02z_ablate_ec2.txt).
#   bash tools/ablate.sh build "1 2 4 8 16"   (here: cross-compile lib/abl/libsgb_abl_<n>.so)
#   bash tools/ablate.sh run   "1 2 4 8 16" <tag>  (on the GPU box)
set -e
L=seggroup_b200/lib; C=seggroup_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default -I include"
case $1 in
build) mkdir -p $L/abl
  for n in $2; do
    for f in edgeconv_bwd_tc; do nvcc $FLAGS -DSGB_ABL=$n -c $C/$f.cu -o $L/abl/${f}_$n.o & done; wait
    objs=$(ls $L/*.o | grep -v -e /edgeconv_bwd_tc.o)
    nvcc -shared -o $L/abl/libsgb_abl_$n.so $objs $L/abl/edgeconv_bwd_tc_$n.o -gencode arch=compute_100a,code=sm_100a -lcudart
    rm $L/abl/*_$n.o
  done ;;
run) for n in 0 $2; do
    echo "== ablation $n"
    if [ $n = 0 ]; then unset SGB_LIB_PATH; else export SGB_LIB_PATH=$PWD/$L/abl/libsgb_abl_$n.so; fi
    timeout 300 python tools/bench_kernels.py --points 150000 --only edgeconv --reps 10 2>&1 | grep MLP3 | cut -c1-150
  done > gpurun_out/$3_ablate.txt 2>&1; cat gpurun_out/$3_ablate.txt ;;
esac
