#!/bin/bash
# Development tool: builds soc_b200/_lib/libsoc_b200_<name>.so with extra nvcc flags (e.g. -DQ_STAGE_N=32);
# run with SOC_B200_LIB=<that file>.   usage: tools/build_variant.sh name [nvcc flags...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=soc_b200/_lib/variant_$name
mkdir -p $out
for f in api sim map sca aux; do
  if [ $f = sim ] || [ ! -f $out/$f.o ]; then
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include -I soc_b200/csrc \
      --ftz=false --prec-div=true --prec-sqrt=true -DSOC_BUILDING "$@" -c soc_b200/csrc/$f.cu -o $out/$f.o &
  fi
done
wait
/usr/local/cuda/bin/nvcc -shared -o soc_b200/_lib/libsoc_b200_$name.so $out/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo soc_b200/_lib/libsoc_b200_$name.so
