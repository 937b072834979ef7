#!/bin/bash
# Compare development builds of libb2l.so (build/variants/libb2l_<name>.so) on the headline bench.
# usage: tools/variant_bench.sh NAME...
for v in "$@"; do
  line=$(B2L_LIB_PATH=$PWD/build/variants/libb2l_$v.so timeout 300 python bench.py --steps 128 --warmup 8 --no-cpu-baseline --headline-only 2>&1 | tail -1)
  echo "$v -> $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "tok/s", round(d["ms_per_step"],4), "ms")' 2>&1 | tail -1)"
done
