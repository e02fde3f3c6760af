set -x
bash tools/ncu_capture.sh r01b_ew_add_contig256 "ew_kernel" 2 python tools/op_sweep.py --filter "single Add" --reps 2
bash tools/ncu_capture.sh r01b_ew_abs_contig256 "ew_kernel" 2 python tools/op_sweep.py --filter "single Abs" --reps 2
bash tools/ncu_capture.sh r01b_fill256 "ew_kernel" 2 python tools/op_sweep.py --filter "single FillConst" --reps 2
bash tools/ncu_capture.sh r01b_bool_and "ew_kernel" 2 python tools/op_sweep.py --filter "bool And" --reps 2
bash tools/ncu_capture.sh r01b_compact_get "compact_kernel" 2 python tools/perf_sweep.py --filter maskedGet --reps 2
bash tools/ncu_capture.sh r01b_compact_trueidx "compact_kernel" 2 python tools/perf_sweep.py --filter trueIdx --reps 2
bash tools/ncu_capture.sh r01b_compact_set "compact_kernel" 2 python tools/perf_sweep.py --filter maskedSet --reps 2
rm -f gpurun_out/r01b_*.source.csv
python tools/op_sweep.py --out gpurun_out/r01b_op_sweep.txt > /dev/null 2>&1
python tools/perf_sweep.py > gpurun_out/r01b_perf_sweep.txt 2>&1
