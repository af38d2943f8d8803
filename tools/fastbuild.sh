#!/bin/bash
# Experimental 2-D float32-only build of the library (development): tools/fastbuild.sh TAG [-DFLAGS...]
# -> tools/bin/libexp_TAG.so ; run with LIBCPAB_B200_SO=tools/bin/libexp_TAG.so
set -e
tag=$1; shift
dim=${CPAB_DIM:-2}
cd "$(dirname "$0")/../libcpab_b200/csrc"
out=../../tools/bin/exp_$tag; mkdir -p $out
for f in cpab_abi cpab_integrate cpab_adjoint_1d cpab_adjoint_2d cpab_adjoint_3d cpab_expm cpab_interp cpab_probe cpab_closed1d cpab_closednd; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       -DCPAB_FAST_BUILD -DCPAB_FAST_DIM=$dim "$@" -c $f.cu -o $out/$f.o -Xptxas -v 2> $out/$f.log &
done
wait
nvcc -shared -o ../../tools/bin/libexp_$tag.so $out/*.o -gencode arch=compute_100a,code=sm_100a
echo built tools/bin/libexp_$tag.so
