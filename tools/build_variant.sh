#!/bin/bash
# tools/build_variant.sh NAME FILE.cu "EXTRA FLAGS" — rebuilds one translation unit with extra flags and links it with
# the other (already built) objects into waterlily_b200/variants/libwl_NAME.so; select it with WL_LIB_PATH.
set -e
cd "$(dirname "$0")/../waterlily_b200/csrc"
mkdir -p ../variants
NAME=$1; FILE=$2; FLAGS=$3
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $FLAGS -c -o ../variants/${NAME}_${FILE%.cu}.o $FILE
OBJS=""
for f in wl_api wl_bc wl_poisson wl_momentum wl_dist wl_vcycle wl_convdiff_tma wl_metrics wl_resjac; do
  if [ "$f.cu" == "$FILE" ]; then OBJS="$OBJS ../variants/${NAME}_$f.o"; else OBJS="$OBJS $f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libwl_$NAME.so $OBJS -lcudart -ldl
echo built ../variants/libwl_$NAME.so
