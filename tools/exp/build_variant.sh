#!/bin/bash
# build_variant.sh NAME -DFOO=1 ... : an A/B build of the library into tools/exp/libs/libtcn_NAME.so (TCN_LIB_PATH selects it)
set -e
name=$1; shift
cd "$(dirname "$0")/../.."
out=tools/exp/libs; mkdir -p $out/obj_$name
for f in computervision_codes_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $f -o $out/obj_$name/$(basename $f).o &
done
wait
nvcc -shared -o $out/libtcn_$name.so $out/obj_$name/*.o
rm -rf $out/obj_$name
echo $out/libtcn_$name.so
