# Round-2 after-state captures (the before-state of each is in profiles/r02_*): the fused Max+ArgMax fold after the
# per-vector arg fold, the sub-word sum through the dot-product unit.
set -x
bash tools/ncu_capture.sh r02b_reduce_max_argmax_c3 "reduce_rows_kernel" 2 python tools/perf_sweep.py --filter "C3 f32 max+argmax" --reps 2
bash tools/ncu_capture.sh r02b_reduce_argmax_c3 "reduce_rows_kernel" 2 python tools/perf_sweep.py --filter "C3 f32 argMaxAxis1" --reps 2
rm -f gpurun_out/r02b_*.source.csv
python tools/ncu_summary.py gpurun_out/r02b_*.raw.csv > gpurun_out/r02b_ncu_summary.txt 2>&1
