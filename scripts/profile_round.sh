#!/bin/bash
# Round profile: bench lines, ncu launch list (same command as the bench), ncu --set full of one whole training step.
# Usage (on the GPU box, from the repo root): bash scripts/profile_round.sh r02     then here: python scripts/summarise_profiles.py r02
set -u
TAG=${1:-r02}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_cfg3.json 2> $OUT/${TAG}_bench_cfg3.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_cfg3.csv $CMD > $OUT/${TAG}_launches.log 2>&1
# one whole step (every kernel of it, in launch order) with the full metric set and source counters
ncu --set full --clock-control none --import-source on -k regex:"conv_|dense_head|pool_nhwc|opt_kernel|nchw_to|splitk" --launch-skip 90 -c 20 \
    -o $OUT/${TAG}_full_step $CMD > $OUT/${TAG}_full_step.log 2>&1
for C in cfg1 cfg2 cfg4 cfg5; do
  python bench.py --workload $C --steps 20 --warmup 3 > $OUT/${TAG}_bench_$C.json 2> $OUT/${TAG}_bench_$C.err
done
python bench.py --workload cfg1 --batch 65536 --steps 20 --warmup 3 --no-cpu > $OUT/${TAG}_bench_cfg1_b65536.json 2> $OUT/${TAG}_bench_cfg1_b65536.err
python bench.py --workload cfg2 --batch 8192 --steps 20 --warmup 3 --no-cpu > $OUT/${TAG}_bench_cfg2_b8192.json 2> $OUT/${TAG}_bench_cfg2_b8192.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --no-graph > $OUT/${TAG}_launches_cfg4.log 2>&1
ls -la $OUT | tail -14
