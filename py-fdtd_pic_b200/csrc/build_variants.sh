#!/bin/bash
# Builds tuning variants of the tile kernel geometry into ../variants/ (experiments only).
set -e
mkdir -p ../variants
build() { # name cells c minblocks kdef
  rm -f pf_tile.o
  make -s pf_tile.o TUNE="-DPF_TILE_CELLS=$2 -DPF_TILE_C=$3 -DPF_TILE_MINBLOCKS=$4 -DPF_TILE_KDEF=$5"
  grep -E "Used|spill" pf_tile.o.ptxas.log | sort | uniq -c | sort -rn | head -3
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$1.so pf_host.o pf_ops.o pf_tile.o pf_pic.o pf_halo.o
}
for v in "$@"; do build $v; done
rm -f pf_tile.o; make -s
