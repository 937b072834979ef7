#!/bin/bash
# Sweep the megakernel's tuning knobs (environment variables read at b2l_finalize / launch) on the headline bench.
# usage: [B2L_LIB_PATH=...] tools/mega_knobs.sh OUT.jsonl "B2L_MEGA_L2AHEAD=16" "B2L_MEGA_L2AHEAD=32 B2L_MEGA_STAGES=8" ...
out=$1; shift
: > "$out"
for kv in "" "$@"; do
  line=$(env $kv timeout 300 python bench.py --steps 128 --warmup 8 --no-cpu-baseline --headline-only 2>/dev/null | tail -1)
  echo "{\"knobs\": \"$kv\", \"line\": $line}" >> "$out"
  echo "$kv -> $(echo "$line" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],1), "tok/s", round(d["ms_per_step"],4), "ms")' 2>/dev/null)"
done
