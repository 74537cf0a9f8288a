#!/bin/bash
# Launch list of the kernels OUTSIDE the NMF iteration loop at the full bench size (512 clips): what the step spends
# on STFT, masks, fill, permutation, iSTFT.  Usage (under gpurun): bash profiles/run_outside_loop.sh <round-tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
ARGS="--steps 1 --warmup 3 --no-cpu-baseline --legs none"
python bench.py $ARGS > $OUT/plain_outside_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k regex:'(stft|istft|gap_mask|range_mask|compact|colsum|colsum_reduce|fill|fill_rows|mean|init_w|init_h|invert_flags|build_perm|gather_rows|scatter_rows|export_state|unpack_factors|unpack_w|transpose_h|finalize|err_reduce)_kernel' \
    --csv --log-file $OUT/launches_outside_${TAG}.csv python bench.py $ARGS > $OUT/ncu_outside_${TAG}.log 2>&1
echo "outside-loop launch list rc=$?"
tail -5 $OUT/ncu_outside_${TAG}.log
