#!/bin/bash
# tools/build_variant.sh NAME FILE.cu "-DFLAG=.. ..."  -> sph-bvf_b200/variants/libsphbvf_NAME.so
# A/B builds while tuning one kernel file: every other object comes from the normal in-tree build.
# Pick a variant at run time with SPHBVF_LIB=<path> (sph-bvf_b200/capi.py).
set -e
cd "$(dirname "$0")/../sph-bvf_b200"
name=$1; file=$2; flags=$3
make -j8 >/dev/null
mkdir -p build/var_$name variants
base=$(basename $file .cu)
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
  -DSPHBVF_WITH_NCCL -Xptxas -v $flags -c csrc/$base.cu -o build/var_$name/$base.o 2> build/var_$name/$base.ptxas.log
objs=""
for o in build/*.o; do
  if [ "$(basename $o)" = "$base.o" ]; then objs="$objs build/var_$name/$base.o"; else objs="$objs $o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libsphbvf_$name.so $objs -lcudart -ldl
echo "built variants/libsphbvf_$name.so"
