#!/bin/bash
# build_full_variant.sh NAME MODEL "EXTRA NVCC FLAGS": like build_variant.sh, but the host translation unit is rebuilt with
# the flags too (for switches the host reads, e.g. -DB2D_RING_CAP_FIXED=14).  Only solve MODEL with the resulting library.
set -e
cd "$(dirname "$0")/../delaydiffeq.jl_b200"
NAME=$1; MODEL=$2; EXTRA=$3
mkdir -p build/var_$NAME lib/variants
JAC=0; case $MODEL in StiffModel|OneDelayModel|SirDdaeModel) JAC=1;; esac
FLAGS="-gencode arch=compute_100a,code=sm_100a -Xfatbin -compress-all -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -I../include -Icsrc $EXTRA"
/usr/local/cuda/bin/nvcc $FLAGS -Xptxas -v -DB2D_INST_MODEL=$MODEL -DB2D_INST_JAC=$JAC -c -o build/var_$NAME/inst_$MODEL.o csrc/inst.cu 2> build/var_$NAME/inst_$MODEL.log &
/usr/local/cuda/bin/nvcc $FLAGS -c -o build/var_$NAME/b200dde.o csrc/b200dde.cu &
wait
OBJS=$(ls build/*.o | grep -v "inst_$MODEL.o" | grep -v "b200dde.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/variants/$NAME.so $OBJS build/var_$NAME/inst_$MODEL.o build/var_$NAME/b200dde.o -ldl
grep -A3 "mos_kernelINS_[0-9]*${MODEL}ELi0ELi0" build/var_$NAME/inst_$MODEL.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $NAME"
