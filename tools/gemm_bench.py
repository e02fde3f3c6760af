"""GEMM throughput on one GPU: config-5 shapes (MLP 784-4096-4096-10, batch 8192) forward and backward."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepnet_b200 import CudaTensor, Tensor, dtypes


def wrap(t):
    return CudaTensor.usingPtr(t.data_ptr(), tuple(t.shape), dtypes.DN_F32, owner=t)


def main():
    dev = CudaTensor.dev()
    dev.Init(0)
    dev.SetStream(torch.cuda.current_stream().cuda_stream)
    # usage: gemm_bench.py [tf32|fp32|strict] [case-name-substring]
    mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
    only = sys.argv[2] if len(sys.argv) > 2 else ""
    dev.SetMathMode(mode)
    print(f"math mode: {mode}")
    torch.backends.cuda.matmul.allow_tf32 = True
    cases = [
        ("fwd1  X[8192,784] . W1.T[784,4096]", 8192, 4096, 784, "NT"),
        ("fwd2  H[8192,4096] . W2.T[4096,4096]", 8192, 4096, 4096, "NT"),
        ("fwd3  H[8192,4096] . W3.T[4096,10]", 8192, 10, 4096, "NT"),
        ("dX2   dY[8192,4096] . W2[4096,4096]", 8192, 4096, 4096, "NN"),
        ("dW2   dY.T[4096,8192] . H[8192,4096]", 4096, 4096, 8192, "TN"),
        ("dW1   dY.T[4096,8192] . X[8192,784]", 4096, 784, 8192, "TN"),
        ("sq    8192^3", 8192, 8192, 8192, "NT"),
    ]
    for name, M, N, K, lay in cases:
        if only and only not in name:
            continue
        if lay == "NT":
            ta, tb = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
            a, b = wrap(ta), wrap(tb).T
            ref = lambda: ta @ tb.T
        elif lay == "NN":
            ta, tb = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
            a, b = wrap(ta), wrap(tb)
            ref = lambda: ta @ tb
        else:
            ta, tb = torch.randn(K, M, device="cuda"), torch.randn(K, N, device="cuda")
            a, b = wrap(ta).T, wrap(tb)
            ref = lambda: ta.T @ tb
        tc = torch.empty(M, N, device="cuda")
        c = wrap(tc)
        fn = lambda: c.FillDot(a, b)
        res = {}
        for label, f in (("ours", fn), ("cublas-tf32", ref)):
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                f()
            e.record()
            e.synchronize()
            res[label] = s.elapsed_time(e) / 10
        err = (tc - ref()).abs().max().item() / (ref().abs().max().item() + 1e-30)
        fl = 2.0 * M * N * K
        print(f"{name:42s} ours {res['ours']:8.3f} ms {fl / res['ours'] / 1e9:8.1f} TFLOP/s | cuBLAS tf32 "
              f"{res['cublas-tf32']:8.3f} ms {fl / res['cublas-tf32'] / 1e9:8.1f} TFLOP/s | max rel diff {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
