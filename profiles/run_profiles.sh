#!/bin/bash
# Runs on the GPU box (under gpurun): the full bench line, then -- each only after the same command has run plainly and
# exited 0 -- the ncu launch list of one bench step and one `--set full` capture of the iteration's kernels, for the clip
# batch (c4, 64 clips) and for the long signal (c5, a 600 s prefix).  Usage: bash profiles/run_profiles.sh <round-tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
C4="--clips 148 --steps 1 --legs none --no-cpu-baseline"
C5="--workload c5 --seconds 600 --steps 1 --no-cpu-baseline"
KERN='(h_step_ts|xht_ts|w_side|w_finish|hbad|hbad_reduce|gather_rows|scatter_rows|build_perm|invert_flags|gram|reduce_splits|reduce_partials|stop|stft|istft|gap_mask|compact|colsum|colsum_reduce|fill|fill_rows|mean|init_w|init_h|finalize|err_reduce|transpose_h|unpack_w|numpy_normals|copy_indices|viol_sum|export_state|status_summary|count_not_done|range_mask|pass_through_all_bad)_kernel'
python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || { echo "bench failed"; tail -20 $OUT/bench_${TAG}.err; exit 1; }
tail -c 600 $OUT/bench_${TAG}.json; echo
python bench.py $C4 > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERN" -s 0 -c 1400 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py $C4 > $OUT/ncu_launches_${TAG}.log 2>&1
echo "c4 launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'h_step_ts_kernel|xht_ts_kernel|w_side_kernel' -s 30 -c 3 -o $OUT/prof_${TAG} \
    python bench.py $C4 > $OUT/ncu_full_${TAG}.log 2>&1
echo "c4 full capture rc=$?"
python bench.py $C5 > $OUT/plain_c5_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERN" -s 0 -c 1400 --csv --log-file $OUT/launches_c5_${TAG}.csv \
    python bench.py $C5 > $OUT/ncu_launches_c5_${TAG}.log 2>&1
echo "c5 launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'h_step_ts_kernel|xht_ts_kernel|w_side_kernel' -s 30 -c 3 -o $OUT/prof_c5_${TAG} \
    python bench.py $C5 > $OUT/ncu_full_c5_${TAG}.log 2>&1
echo "c5 full capture rc=$?"
ls -la $OUT | tail -20
