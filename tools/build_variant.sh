#!/bin/bash
# builds a variant of the library with extra nvcc defines into idelucs_b200/lib/var_<name>/ (select it with IDELUCS_B200_LIB)
# usage: tools/build_variant.sh <name> -DFOO=1 ...
set -e
name=$1; shift
d=idelucs_b200/lib/var_$name
mkdir -p $d
for f in kernels iid_loss fasta train_ops; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -diag-suppress 177 "$@" -c idelucs_b200/csrc/$f.cu -o $d/$f.o &
done
wait
nvcc -shared -o $d/libidelucs_b200.so $d/*.o -gencode arch=compute_100a,code=sm_100a
echo built $d/libidelucs_b200.so
