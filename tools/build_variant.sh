#!/bin/bash
# tools/build_variant.sh NAME "EXTRA NVCC FLAGS" [files...]: experiment build of liboctcube_b200 -> octcubem_b200/variants/libNAME.so
# Only the listed sources (default: the two attention kernels) are recompiled with the flags; the rest is linked from csrc/build.
set -e
cd "$(dirname "$0")/../octcubem_b200/csrc"
name=$1; extra=$2; shift 2 || true
files=${@:-"attn_tc attn_bwd_tc"}
make -j8 >/dev/null
mkdir -p build_$name ../variants
objs=""
for f in api mask pool ingest ln loss gemm_simt gemm_tc attn_simt attn_tc attn_bwd_tc patch_embed_tc optim clip allreduce dispatch; do
  if [[ " $files " == *" $f "* ]]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
      -Xptxas -v -DOCT_BUILDING $extra -c $f.cu -o build_$name/$f.o 2> build_$name/$f.ptxas.log || { cat build_$name/$f.ptxas.log; exit 1; }
    objs="$objs build_$name/$f.o"
  else
    objs="$objs build/$f.o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib$name.so $objs -cudart static
echo "built variants/lib$name.so ($extra)"
