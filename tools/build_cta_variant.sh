#!/bin/bash
# build_cta_variant.sh NAME MODEL THREADS MINB: library variant with another CTA size for MODEL's kernels (the host
# translation unit is rebuilt too: ring slabs and occupancy are computed from the CTA size).  Other models keep
# their regular objects, so only solve MODEL with this library.  Output: lib/variants/NAME.so
set -e
cd "$(dirname "$0")/../delaydiffeq.jl_b200"
NAME=$1; MODEL=$2; THREADS=$3; MINB=$4
mkdir -p build/var_$NAME lib/variants
JAC=0; case $MODEL in StiffModel|OneDelayModel|SirDdaeModel) JAC=1;; esac
FLAGS="-gencode arch=compute_100a,code=sm_100a -Xfatbin -compress-all -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -I../include -Icsrc -DB2D_CTA_THREADS=$THREADS -DB2D_MINB_OVERRIDE=$MINB"
/usr/local/cuda/bin/nvcc $FLAGS -Xptxas -v -DB2D_INST_MODEL=$MODEL -DB2D_INST_JAC=$JAC -c -o build/var_$NAME/inst_$MODEL.o csrc/inst.cu 2> build/var_$NAME/inst_$MODEL.log &
/usr/local/cuda/bin/nvcc $FLAGS -c -o build/var_$NAME/b200dde.o csrc/b200dde.cu &
wait
OBJS=$(ls build/*.o | grep -v "inst_$MODEL.o" | grep -v "b200dde.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/variants/$NAME.so $OBJS build/var_$NAME/inst_$MODEL.o build/var_$NAME/b200dde.o -ldl
grep -A3 "mos_kernelINS_[0-9]*${MODEL}ELi0ELi0" build/var_$NAME/inst_$MODEL.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $NAME"
