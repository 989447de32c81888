#!/bin/bash
# The profile set of a round, to be run on the GPU box (one GPU):  gpurun --timeout 1500 -- 'bash tools/profile_round.sh r2a'
# Writes gpurun_out/<tag>_*; read the reports back with `ncu -i <rep> --page raw --csv` and profiles/func_profile.py /
# profiles/line_profile.py, then copy the summaries into profiles/.
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 16 --warmup 3"
# launch list of the whole step (serialised, cold caches: compare SHARES with bench.py's share_alone, not absolute times)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_launches.log 2>&1
# one full capture per kernel family
ncu --set full --clock-control none --import-source on -k regex:me_sb_kernel  -s 3 -c 1 -o $OUT/${TAG}_me  $B > $OUT/${TAG}_me.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:txq_kernel    -s 9 -c 1 -o $OUT/${TAG}_txq $B > $OUT/${TAG}_txq.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lf_row_kernel -s 1 -c 1 -o $OUT/${TAG}_lf  $B > $OUT/${TAG}_lf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pred_kernel   -s 20 -c 1 -o $OUT/${TAG}_pred python tools/pred_times.py > $OUT/${TAG}_pred.log 2>&1
# un-profiled numbers of the same state
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python tools/lf_times.py > $OUT/${TAG}_lf_times.txt 2>&1
python tools/pred_times.py > $OUT/${TAG}_pred_times.txt 2>&1
ls -la $OUT/${TAG}_*
