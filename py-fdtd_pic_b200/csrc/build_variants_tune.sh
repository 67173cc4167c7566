#!/bin/bash
# Builds pf_tile.cu with arbitrary -D tuning flags into ../variants/lib_<name>.so (experiments only).
# usage: ./build_variants_tune.sh "name -DFOO=1 -DBAR" ...
set -e
mkdir -p ../variants
for v in "$@"; do
  set -- $v; name=$1; shift
  rm -f pf_tile.o
  make -s pf_tile.o TUNE="$*"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so pf_host.o pf_ops.o pf_tile.o pf_pic.o pf_halo.o
  echo "$name: $(grep -c spill pf_tile.o.ptxas.log) kernels"
done
rm -f pf_tile.o; make -s
