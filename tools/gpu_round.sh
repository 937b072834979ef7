#!/bin/bash
# One gpurun call that refreshes every artefact under profiles/ for the current build:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
# 1. GPU parity suite, 2. default bench line, 3. ncu launch list of the headline bench, 4. ncu --set full of one
# 4-token megakernel launch (+ raw/ source pages as csv), 5. phase timeline.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/gputest.log 2>&1
  echo "pytest exit $?" >> $O/gputest.log
fi
timeout 600 python bench.py > $O/bench_line.json 2> $O/bench.err
echo "bench exit $?" >> $O/bench.err
timeout 300 python bench.py --headline-only --no-cpu-baseline --steps 32 --warmup 3 > $O/bench_short.json 2>> $O/bench.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --headline-only --no-cpu-baseline --steps 32 --warmup 3 > $O/ncu_launches.log 2>&1
timeout 120 python tools/mega_ncu.py 4 > $O/mega_ncu_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mega_decode_kernel -s 1 -c 1 -f -o $O/mega_full \
    python tools/mega_ncu.py 4 > $O/ncu_full.log 2>&1
if [ -f $O/mega_full.ncu-rep ]; then
  ncu -i $O/mega_full.ncu-rep --page raw --csv > $O/mega_full_raw.csv 2>/dev/null
  ncu -i $O/mega_full.ncu-rep --page details > $O/mega_full_details.txt 2>/dev/null
  ncu -i $O/mega_full.ncu-rep --page source --csv > $O/mega_full_source.csv 2>/dev/null
  [ "${KEEP_REP:-0}" = "1" ] || rm -f $O/mega_full.ncu-rep
fi
[ -n "${EXTRA:-}" ] && bash -c "$EXTRA" > $O/extra.log 2>&1
ls -la $O > $O/listing.txt
tail -3 $O/gputest.log 2>/dev/null
cat $O/bench_line.json | cut -c1-600
