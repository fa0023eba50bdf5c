#!/bin/bash
# Builds the CUDA library of a git ref (or of the working tree with NAME=wt) into real-time-voice-cloning_b200/_variants/NAME.so
# so two kernels can be timed against each other in ONE gpurun call (select with WRNN_B200_LIB=<path>).
set -e
cd "$(dirname "$0")/.."
NAME=$1; REF=${2:-}
PKG=real-time-voice-cloning_b200
OUT=$PKG/_variants; mkdir -p $OUT/$NAME
if [ -n "$REF" ]; then git archive $REF $PKG/csrc include | tar -x -C $OUT/$NAME; else mkdir -p $OUT/$NAME/$PKG; cp -r $PKG/csrc $OUT/$NAME/$PKG/; cp -r include $OUT/$NAME/; rm -f $OUT/$NAME/$PKG/csrc/*.o; fi
cd $OUT/$NAME/$PKG/csrc
objs=""
for f in *.cu; do nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -O2 $EXTRA -c $f -o ${f%.cu}.o & done; wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../../$NAME.so *.o
cd ../../..; rm -rf $NAME; ls -la $NAME.so
