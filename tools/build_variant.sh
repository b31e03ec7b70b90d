#!/bin/bash
# usage: tools/build_variant.sh <name> <csrc dir>   -> build/variants/lib_<name>.so  (kernel A/B runs: MUSE_B200_LIB=...)
set -e
d=/tmp/muse_variant_$1
rm -rf $d && mkdir -p $d/go-muse_b200 $d/include build/variants
cp -r $2 $d/go-muse_b200/csrc && cp include/muse_b200.h $d/include/
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -o build/variants/lib_$1.so $d/go-muse_b200/csrc/muse_api.cu 2>&1 | grep -E "error" || true
ls -la build/variants/lib_$1.so
