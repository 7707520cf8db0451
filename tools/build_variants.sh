#!/bin/bash
# Builds variants of a kernel file for A/B timing on the GPU box: only $FILES (default: traverse; e.g. FILES="pt vcm") are recompiled
# with the extra -D flags, the other objects come from the regular build.  usage: tools/build_variants.sh name1:"-DX=1 -DY=2" name2:"..."
# Outputs imbatracer_b200/libimba_var_<name>.so (git-ignored); run with IMBA_LIB_VARIANT=libimba_var_<name>.so tools/bench_quick.py
set -e
cd "$(dirname "$0")/../imbatracer_b200/csrc"
make -j8 > /dev/null
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  mkdir -p build/var_$name
  ( objs=""; others=$(ls build/*.o)
    for f in ${FILES:-traverse}; do
      /usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
        -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v $flags -c -o build/var_$name/$f.o $f.cu 2> build/var_$name/$f.ptxas.log || exit 1
      objs="$objs build/var_$name/$f.o"; others=$(echo "$others" | grep -v "/$f.o")
    done
    /usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -Xlinker --no-undefined \
      -o ../libimba_var_$name.so $objs $others -lcuda && echo "built $name ($flags)" ) &
done
wait
