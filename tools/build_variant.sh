#!/bin/bash
# build_variant.sh NAME MODEL "EXTRA NVCC FLAGS": a library variant for kernel experiments — only MODEL's translation
# unit is recompiled with the extra flags, every other object comes from the regular build.  Output:
# delaydiffeq.jl_b200/lib/variants/NAME.so (travels to the GPU box; select it with B200DDE_LIB=...).
set -e
cd "$(dirname "$0")/../delaydiffeq.jl_b200"
NAME=$1; MODEL=$2; EXTRA=$3
mkdir -p build/var_$NAME lib/variants
JAC=0; case $MODEL in StiffModel|OneDelayModel|SirDdaeModel) JAC=1;; esac
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -Xfatbin -compress-all -O3 -std=c++17 -lineinfo -fmad=false \
  -Xcompiler -fPIC -I../include -Icsrc -Xptxas -v $EXTRA -DB2D_INST_MODEL=$MODEL -DB2D_INST_JAC=$JAC -c -o build/var_$NAME/inst_$MODEL.o csrc/inst.cu 2> build/var_$NAME/inst_$MODEL.log
OBJS=$(ls build/*.o | grep -v "inst_$MODEL.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/variants/$NAME.so $OBJS build/var_$NAME/inst_$MODEL.o -ldl
grep -A3 "mos_kernelINS_[0-9]*${MODEL}ELi0ELi0" build/var_$NAME/inst_$MODEL.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $NAME"
