#!/bin/bash
# build nalo_slam_b200/libnalo_gpu_alt<suffix>.so with extra nvcc flags (A/B experiments): tools/build_alt.sh <suffix> <flags...>
set -e
suf=$1; shift
cd "$(dirname "$0")/../nalo_slam_b200/csrc"
out=/tmp/altbuild$suf; mkdir -p $out
for f in *.cu; do /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include "$@" -c -o $out/${f%.cu}.o $f & done; wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o ../libnalo_gpu_alt$suf.so $out/*.o
ls -la ../libnalo_gpu_alt$suf.so
