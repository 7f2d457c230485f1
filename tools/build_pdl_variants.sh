#!/bin/bash
# Variants of the programmatic-dependent-launch trigger policy (-DSONIC_PDL_TRIGGER=0 / 2; the stock library is 1):
# rebuilds the three sources that contain PDL kernels and links variants/libsonic_pdlt<N>.so (select with SONIC_LIB).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/sonicdiffusionbayeslab_b200/csrc
OUT=$ROOT/variants
mkdir -p $OUT/obj
make -C $CS -j8 > /dev/null
ARCH="-gencode arch=compute_100a,code=sm_100a"
for T in "$@"; do
  OBJS=""
  for o in $CS/build/*.o; do
    b=$(basename $o .o)
    if [ "$b" == gemm ] || [ "$b" == attention ] || [ "$b" == norm ]; then
      nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC --expt-relaxed-constexpr -DSONIC_PDL_TRIGGER=$T -c $CS/$b.cu -o $OUT/obj/pdlt${T}_$b.o &
      OBJS="$OBJS $OUT/obj/pdlt${T}_$b.o"
    else
      OBJS="$OBJS $o"
    fi
  done
  wait
  nvcc $ARCH -shared -o $OUT/libsonic_pdlt$T.so $OBJS -cudart shared
  echo built $OUT/libsonic_pdlt$T.so
done
