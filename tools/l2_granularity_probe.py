"""Experiment: effect of cudaLimitMaxL2FetchGranularity (32/64/128 B) on gather/scatter and on streaming kernels."""
import ctypes
import os
import subprocess
import sys

for gran in (128, 64, 32):
    code = f"""
import ctypes, sys, os
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import torch
torch.zeros(1, device='cuda')
rt = ctypes.CDLL('libcudart.so.12')
cudaLimitMaxL2FetchGranularity = 0x05
print('set', rt.cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, ctypes.c_size_t({gran})), flush=True)
v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), cudaLimitMaxL2FetchGranularity); print('granularity now', v.value, flush=True)
sys.argv = ['perf_sweep', '--filter', sys.argv[1], '--reps', '3']
sys.path.insert(0, os.path.join({os.path.dirname(os.path.abspath(__file__))!r}))
import perf_sweep; perf_sweep.main()
"""
    for flt in ("C4 i64", "single add contiguous", "single add a.T", "C3 f32 sum"):
        out = subprocess.run([sys.executable, "-c", code, flt], capture_output=True, text=True)
        for line in out.stdout.splitlines():
            if "GB/s" in line or "granularity now" in line:
                print(gran, line)
        if out.returncode:
            print(out.stderr[-500:])
