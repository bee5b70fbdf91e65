#!/bin/bash
# Round profile: bench line, ncu launch list (same command), ncu --set full of the heaviest kernels.
# Usage (on the GPU box, from the repo root): bash scripts/profile_round.sh r01
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-graph"
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_cfg3.json 2> $OUT/${TAG}_bench_cfg3.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_cfg3.csv $CMD > $OUT/${TAG}_launches.log 2>&1
for K in conv_c3_fwd_kernel conv_c3_df_kernel conv_df_kernel conv_igemm_kernel pool_nhwc_bwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip 4 -c 1 -o $OUT/${TAG}_full_$K $CMD > $OUT/${TAG}_full_$K.log 2>&1
done
ls -la $OUT | tail -12
# the other configs (bench lines only) and the LSTM launch list
for C in cfg1 cfg2 cfg4 cfg5; do
  python bench.py --workload $C --steps 20 --warmup 3 > $OUT/${TAG}_bench_$C.json 2> $OUT/${TAG}_bench_$C.err
done
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu --no-graph > $OUT/${TAG}_launches_cfg4.log 2>&1
ls -la $OUT | tail -12
