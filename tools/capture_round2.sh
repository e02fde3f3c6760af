# Round-2 ncu evidence (run on the GPU box through gpurun; CSV summaries land in gpurun_out/, the ones worth keeping
# are copied to profiles/). One kernel per capture, `--set full --clock-control none --import-source on`.
set -x
# launch list of the headline step (shares of the step per kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-verify > gpurun_out/r02_bench_under_ncu.log 2>&1
bash tools/ncu_capture.sh r02_ew_add_contig256_f64 "ew_kernel" 0 python tools/perf_sweep.py --filter "C2 double add contiguous" --reps 2
bash tools/ncu_capture.sh r02_fused_spec_tanh_grad "ew_kernel" 2 python tools/perf_sweep.py --filter "FUSED f32 dh" --reps 2
bash tools/ncu_capture.sh r02_reduce_max_argmax_c3 "reduce_rows_kernel" 2 python tools/perf_sweep.py --filter "C3 f32 max+argmax" --reps 2
bash tools/ncu_capture.sh r02_gemm_tf32_2cta_8192x4096x4096 "gemm_tf32_2cta_kernel" 2 python tools/gemm_bench.py tf32 fwd2
bash tools/ncu_capture.sh r02_gemm_3xtf32_2cta_8192x4096x4096 "gemm_tf32_2cta_kernel" 2 python tools/gemm_bench.py fp32 fwd2
bash tools/ncu_capture.sh r02_scatter_partition "scatter_partition_kernel" 2 python tools/perf_sweep.py --filter "C4 i64 scatter random" --reps 2
bash tools/ncu_capture.sh r02_scatter_accumulate "scatter_accumulate_kernel" 1 python tools/perf_sweep.py --filter "C4 i64 scatter random" --reps 2
rm -f gpurun_out/r02_*.source.csv
python tools/ncu_summary.py gpurun_out/r02_*.raw.csv > gpurun_out/r02_ncu_summary.txt 2>&1
