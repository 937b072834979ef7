#!/bin/bash
# Development build of libb2l.so into build/variants/libb2l_<name>.so (selected at run time with B2L_LIB_PATH, see
# tools/variant_bench.sh). usage: tools/build_variant.sh NAME [-DFLAG ...]   (-DMEGA_ONLY_1B compiles only the 1B instantiations)
set -e
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" \
  -shared -o build/variants/libb2l_$name.so gabby_b200/csrc/engine.cu -ldl 2>&1 | grep -E "error|warning: v|ptxas info" | head -20
ls -la build/variants/libb2l_$name.so
