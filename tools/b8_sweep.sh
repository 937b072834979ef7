#!/bin/bash
# A/B runs of the batch-8 decode step (bench.py --workload 3b-b8) under env switches; one line per variant.
# usage: tools/b8_sweep.sh "VAR=val VAR2=val" "..." ...   ("" = defaults)
for v in "$@"; do
  line=$(env $v timeout 200 python bench.py --workload 3b-b8 --steps 32 --warmup 4 --no-cpu-baseline --headline-only 2>&1 | tail -1)
  echo "[$v] -> $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "tok/s", round(d["ms_per_step"],4), "ms/step", d["roofline"]["kernel"])' 2>&1 | tail -1)"
done
