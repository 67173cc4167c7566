#!/bin/bash
# tools/build_variant.sh <name> "<-D flags>" [file.cu ...]  -> py-fdtd_pic_b200/variants/lib_<name>.so
# Rebuilds only the listed sources (default pf_tile.cu) with the extra flags and links them with the default
# objects of the other sources; select the result with PYFDTD_B200_LIB=... (kernel tuning A/B on the GPU box).
set -e
name=$1; flags=$2; shift 2 || true
files=${@:-pf_tile.cu}
cd "$(dirname "$0")/../py-fdtd_pic_b200/csrc"
make -s -j4 >/dev/null
mkdir -p ../variants /tmp/pfv_$name
objs=""
for f in pf_host.cu pf_setup.cu pf_probe.cu pf_ops.cu pf_tile.cu pf_pic.cu pf_halo.cu pf_dormant.cu; do
  if [[ " $files " == *" $f "* ]]; then
    nvcc $flags -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -ffp-contract=off -Xptxas -v --fmad=false -c $f -o /tmp/pfv_$name/${f%.cu}.o 2> /tmp/pfv_$name/${f%.cu}.log
    objs="$objs /tmp/pfv_$name/${f%.cu}.o"
  else
    objs="$objs ${f%.cu}.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so $objs
echo "built variants/lib_$name.so"
