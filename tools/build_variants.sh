#!/bin/bash
# Development helper: libparasuite_b200 variants that differ in -D flags of pileup.cu / profile.cu, built side by side
# under para-suite_b200/lib/ (git-ignored, travel to the GPU box).  usage: tools/build_variants.sh name "-DX=1 ..." [name flags]...
set -e
R=$(cd "$(dirname "$0")/.." && pwd); C=$R/para-suite_b200/csrc; O=${TMPDIR:-/tmp}/vb; mkdir -p $O
NV="nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3,-pthread -I $R/include -I $C"
for f in ctx; do [ $O/$f.o -nt $C/$f.cu -a $O/$f.o -nt $C/internal.h -a $O/$f.o -nt $R/include/parasuite_b200.h ] || $NV -c $C/$f.cu -o $O/$f.o; done
for f in bam_batcher flush clust_writer tool_loops liftover profile_writer; do [ $O/$f.o -nt $C/$f.cpp -a $O/$f.o -nt $R/include/parasuite_b200.h ] || g++ -O3 -std=c++17 -fPIC -pthread -I $R/include -I $C -I /usr/local/cuda/include -c $C/$f.cpp -o $O/$f.o; done
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  $NV $flags -c $C/pileup.cu -o $O/pileup_$name.o &
  $NV $flags -c $C/profile.cu -o $O/profile_$name.o &
  wait
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $R/para-suite_b200/lib/libparasuite_b200_$name.so $O/ctx.o $O/pileup_$name.o $O/profile_$name.o $O/bam_batcher.o $O/flush.o $O/clust_writer.o $O/tool_loops.o $O/liftover.o $O/profile_writer.o -lz -lpthread
  echo built libparasuite_b200_$name.so
done
