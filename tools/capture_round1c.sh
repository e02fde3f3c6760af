set -x
bash tools/ncu_capture.sh r01c_fused_tanh_grad "ew_kernel" 2 python tools/perf_sweep.py --filter "FUSED f32 dh" --reps 2
bash tools/ncu_capture.sh r01c_compact_trueidx "compact_kernel" 2 python tools/perf_sweep.py --filter "trueIdx [8192,8192] p=0.5" --reps 2
bash tools/ncu_capture.sh r01c_compact_get "compact_kernel" 2 python tools/perf_sweep.py --filter "maskedGet p=0.5" --reps 2
bash tools/ncu_capture.sh r01c_scatter_hotspot "scatter_kernel" 2 python tools/perf_sweep.py --filter "scatter hot-spot" --reps 2
rm -f gpurun_out/r01c_*.source.csv
