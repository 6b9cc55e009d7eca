# compute-sanitizer evidence for the round (run on the GPU box): memcheck over the smoke + a parity subset, racecheck over the
# shared-memory kernels of the reduce family.  (racecheck does not model tcgen05.alloc's asynchronous write of the TMEM address
# into shared memory nor mbarrier-completed bulk copies: the GEMM and the bulk-copy reduce are covered by memcheck + parity.)
mkdir -p gpurun_out
OUT=gpurun_out/${TAG:-r02b}_compute_sanitizer.txt
echo "# compute-sanitizer (CUDA 12.9) on a B200, round 2 (second session: hybrid f32 schedule, scale-copy thread)" > $OUT
run() { # title, command...
  echo "## $1" >> $OUT; shift
  echo "\$ $*" >> $OUT
  timeout 900 "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard|Race reported|access at" | head -24 >> $OUT
}
run smoke_plain python -c "import __graft_entry__ as g; g.smoke()"
run memcheck_smoke compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()"
[ -n "$ONLY_MATMUL" ] || run memcheck_reduce compute-sanitizer --tool memcheck python -m pytest tests/test_reduce_gpu.py -q -m gpu -k "kat or ties or pitched or offset_views or plane_sum or special_values or segmented or back_to_back or integer_pattern"
run memcheck_matmul compute-sanitizer --tool memcheck python -m pytest tests/test_matmul_gpu.py tests/test_matmul_scaled_gpu.py -q -m gpu -k "golden or ragged or int8_exact or mixed or unaligned_row_pitch or tail_split_all_tiles or batched_and or nonfinite or (parity_scaled and 2sm_n224) or (parity_nvfp4 and 300) or prepacked or split_k_tail_on_scaled"
run synccheck_scaled compute-sanitizer --tool synccheck python -m pytest tests/test_matmul_scaled_gpu.py tests/test_matmul_gpu.py -q -m gpu -k "(parity_scaled and 300) or (parity_nvfp4 and 300) or split_modes_batched"
[ -n "$ONLY_MATMUL" ] || B200_REDUCE_VARIANT=u8 run racecheck_reduce_plain compute-sanitizer --tool racecheck python -m pytest tests/test_reduce_gpu.py -q -m gpu -k "kat or ties or pitched or plane_sum or special_values or axis_reductions_16bit"
[ -n "$ONLY_MATMUL" ] || run racecheck_reduce_bulk compute-sanitizer --tool racecheck python -m pytest tests/test_reduce_gpu.py -q -m gpu -k "integer_pattern and tma and 1048581"
cat $OUT
