OUT=gpurun_out; TAG=r02h
C4="--clips 148 --steps 1 --legs none --no-cpu-baseline"
KERN='(h_step_ts|xht_ts|w_side|w_finish|hbad|hbad_reduce|gather_rows|scatter_rows|build_perm|invert_flags|gram|reduce_splits|reduce_partials|stop|stft|istft|gap_mask|compact|colsum|colsum_reduce|fill|fill_rows|mean|init_w|init_h|finalize|err_reduce|transpose_h|unpack_w|numpy_normals|copy_indices|viol_sum|export_state|status_summary|count_not_done|range_mask|pass_through_all_bad|pcm_mono|pcm_normalise|pcm_store)_kernel'
python bench.py $C4 > $OUT/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERN" -s 0 -c 1400 --csv --log-file $OUT/launches_${TAG}.csv python bench.py $C4 > $OUT/ncu_launches_${TAG}.log 2>&1
echo "c4 launch list rc=$?"
