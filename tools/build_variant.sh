#!/bin/bash
# usage: tools/build_variant.sh NAME REV ["NVCC_EXTRA"]  -- compile the CUDA core as it was at git revision REV (or WORK for
# the working tree) into rayito_b200/csrc/_ab/lib_NAME.so, for same-session A/Bs against the current host libraries
name=$1; rev=$2; extra=$3
tmp=$(mktemp -d)
if [ "$rev" = "WORK" ]; then
  cp -r rayito_b200/csrc include $tmp/
else
  git archive $rev rayito_b200/csrc include | tar -x -C $tmp && mv $tmp/rayito_b200/csrc $tmp/csrc
fi
mkdir -p rayito_b200/csrc/_ab
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -ccbin /usr/bin/g++ -Xcompiler -fPIC -shared $extra -I$tmp/include $tmp/csrc/rt_core.cu -o rayito_b200/csrc/_ab/lib_$name.so || exit 1
rm -rf $tmp
echo "built _ab/lib_$name.so from $rev [$extra]"
