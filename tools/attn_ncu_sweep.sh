#!/bin/bash
# ncu durations of the batched decode attention kernel (3B, batch 8, context 2048) under env switches: average of 20 launches.
# (the sweeps committed in profiles/r02_attn_decode_sweeps.txt were made with a version that printed the block size in the grid column)
# usage: tools/attn_ncu_sweep.sh "VAR=val ..." ...
for v in "$@"; do
  env $v timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_decode_mma -s 30 -c 20 --csv --log-file /tmp/attn_sweep.csv \
      python tools/b8_probe.py 3b 8 2048 > /tmp/attn_sweep.log 2>&1
  python - "$v" <<'P'
import csv, sys
vals, grid = [], ""
for r in csv.reader(open("/tmp/attn_sweep.csv")):
    if len(r) > 5 and r[-1].replace(".", "").replace(",", "").isdigit() and "attn_decode" in ",".join(r):
        vals.append(float(r[-1].replace(",", ""))); grid = [x for x in r if x.startswith("(")][-1]   # columns: ..., Block Size, Grid Size, ...
print(f"[{sys.argv[1]}] grid {grid}: attention kernel {sum(vals) / max(1, len(vals)) / 1e3:.2f} us avg of {len(vals)} (min {min(vals or [0]) / 1e3:.2f})")
P
done
