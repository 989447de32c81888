#!/bin/bash
# The profile set of a round, to be run on the GPU box (one GPU):  gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2q'
# Writes gpurun_out/<tag>_*; read the reports back with `ncu -i <rep> --page raw --csv` and profiles/func_profile.py /
# profiles/line_profile.py, then copy the summaries into profiles/.
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --no-encoder --no-cpu --steps 2 --warmup 1"
# launch list of the whole step (serialised, cold caches: compare SHARES with bench.py's share_alone, not absolute times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
# one full capture per kernel family (the 4th..6th launch of each: warm-up pictures skipped)
for k in me_sb_kernel:3 txq_multi_kernel:3 pred_kernel:3 rate_multi_kernel:3 pack_kernel:3 lf_row_kernel:1; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:^$name -s $skip -c 1 -o $OUT/${TAG}_$name $B > $OUT/${TAG}_$name.log 2>&1
  ncu -i $OUT/${TAG}_$name.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
done
# un-profiled numbers of the same state
python bench.py --no-encoder > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
ls -la $OUT/${TAG}_*
