#!/bin/bash
# Runs on the GPU box (under gpurun): full-size bench line, then the ncu launch list of a bench step and one
# `--set full` capture of the three kernels of the iteration.  Usage: bash profiles/run_profiles.sh <round-tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
SMALL="--clips 64 --steps 1 --no-cpu-baseline"
python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || { echo "bench failed"; tail -20 $OUT/bench_${TAG}.err; exit 1; }
tail -c 2500 $OUT/bench_${TAG}.json
python bench.py $SMALL > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'(h_step_ts|xht_ts|w_side|w_finish|hbad|hbad_reduce|gather_rows|scatter_rows|build_perm|invert_flags|gram|reduce_splits|stop|stft|istft|gap_mask|compact|colsum|colsum_reduce|fill|fill_rows|mean|init_w|init_h|finalize|err_reduce|transpose_h|unpack_w)_kernel' -s 0 -c 900 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py $SMALL > $OUT/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
python bench.py $SMALL > $OUT/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'h_step_ts_kernel|xht_ts_kernel|w_side_kernel' -s 30 -c 3 -o $OUT/prof_${TAG} \
    python bench.py $SMALL > $OUT/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la $OUT
