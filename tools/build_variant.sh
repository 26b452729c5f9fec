#!/bin/bash
# tools/build_variant.sh NAME "EXTRA FLAGS" FILE.cu [FILE2.cu …] — rebuilds the named translation units with extra flags and
# links them with the other (already built) objects into waterlily_b200/variants/libwl_NAME.so; select it with WL_LIB_PATH.
# (Flags that change inline functions of the shared headers — WL_AHEAD, WL_PF — must name every unit that uses them:
# wl_poisson.cu wl_vcycle.cu.)
set -e
cd "$(dirname "$0")/../waterlily_b200/csrc"
mkdir -p ../variants
NAME=$1; FLAGS=$2; shift 2
OBJS=""
for f in wl_api wl_bc wl_poisson wl_momentum wl_dist wl_vcycle wl_convdiff_tma wl_metrics wl_resjac; do
  hit=0
  for FILE in "$@"; do if [ "$f.cu" == "$FILE" ]; then hit=1; fi; done
  if [ $hit == 1 ]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $FLAGS -c -o ../variants/${NAME}_$f.o $f.cu
    OBJS="$OBJS ../variants/${NAME}_$f.o"
  else OBJS="$OBJS $f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libwl_$NAME.so $OBJS -lcudart -ldl
echo built ../variants/libwl_$NAME.so
