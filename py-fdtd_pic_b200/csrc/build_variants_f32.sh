#!/bin/bash
# Builds tuning variants of the fp32 tile kernel geometry into ../variants/ (experiments only).
set -e
mkdir -p ../variants
build() { # name c minblocks
  rm -f pf_tile.o
  make -s pf_tile.o TUNE="-DPF_TILE_C_F32=$2 -DPF_TILE_MINBLOCKS_F32=$3"
  grep -A2 Fast32 pf_tile.o.ptxas.log | grep -E "Used|spill" | sort | uniq -c | sort -rn | head -4
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$1.so pf_host.o pf_ops.o pf_tile.o pf_pic.o pf_halo.o
}
for v in "$@"; do build $v; done
rm -f pf_tile.o; make -s
