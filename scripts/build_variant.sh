#!/bin/bash
# Build a variant of the library with extra nvcc flags into build/<name>.so (A/B experiments: PCL_LIB=build/<name>.so).
# usage: scripts/build_variant.sh name "-DFOO=1 -DBAR"
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
mkdir -p $tmp/piccolo_b200 $tmp/include $root/build
cp -r $root/piccolo_b200/csrc $tmp/piccolo_b200/csrc
cp $root/include/*.h $tmp/include/
rm -f $tmp/piccolo_b200/csrc/*.o
make -C $tmp/piccolo_b200/csrc -j16 NVCCFLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $flags" > $tmp/build.log 2>&1 || { tail -20 $tmp/build.log; exit 1; }
cp $tmp/piccolo_b200/libpiccolo_b200.so $root/build/$name.so
rm -rf $tmp
echo built build/$name.so
