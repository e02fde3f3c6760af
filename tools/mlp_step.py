"""Config 5 (BASELINE.json): LearnMnist-style MLP training step, 784-4096-4096-10, batch 8192, synthetic data,
written against the Tensor frontend only — so the same function runs on CudaTensor.Dev (tcgen05 MatMatDot +
element-wise + reduction kernels) and on the HostTensor oracle.

Model math from the reference: layer = input .* weights.T + bias (ML/MLModels/Neural.fs:118-130), tanh hidden
layers, softmax = exp(x - max_1) / sum_1 (ML/MLModels/Util.fs:53-56), loss = mean(-sum_1(target * log pred))
(ML/MLModels/Neural.fs:42-43), weights ~ U(-r, r), r = 4*sqrt(6/(fanIn+fanOut)), zero bias (Neural.fs:92-100),
plain gradient descent pars - step*grad (ML/MLOptimizers/GradientDescent.fs:56-58). The reference derives the
gradient symbolically (Symbolic/SymTensor/Deriv.fs, out of scope); here it is written out by hand."""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from deepnet_b200 import Tensor, dtypes


def init_params(rng: np.random.Generator, sizes, dtype=np.float32):
    """[(W [out, in], b [out])], Neural.fs:92-100."""
    params = []
    for n_in, n_out in zip(sizes[:-1], sizes[1:]):
        r = 4.0 * np.sqrt(6.0 / (n_in + n_out))
        params.append((rng.uniform(-r, r, size=(n_out, n_in)).astype(dtype), np.zeros(n_out, dtype=dtype)))
    return params


def synthetic_batch(rng: np.random.Generator, batch: int, n_in: int, n_class: int, dtype=np.float32):
    x = rng.uniform(0.0, 1.0, size=(batch, n_in)).astype(dtype)
    t = np.zeros((batch, n_class), dtype=dtype)
    t[np.arange(batch), rng.integers(0, n_class, size=batch)] = 1.0
    return x, t


def train_step(x: Tensor, target: Tensor, params, step: float, fused: bool = False):
    """One forward + backward + SGD update. `params` is a list of (W, b) Tensors updated in place.
    Returns (loss tensor (rank 0), predictions). `fused=True` evaluates every chain of element-wise operators with
    one FusedElemwise call (dn_fused_elemwise) — bit-identical results, fewer passes over HBM."""
    if fused:
        return _train_step_fused(x, target, params, step)
    batch = x.Shape[0]
    # forward
    acts = [x]
    h = x
    for li, (w, b) in enumerate(params):
        z = h @ w.T + b                       # input .* weights.T + bias
        if li < len(params) - 1:
            h = z.tanh()
        else:
            c = z.maxAxis(1).padRight()       # softmax, Util.fs:53-56
            y = (z - c).exp()
            h = y / y.sumAxis(1).padRight()
        acts.append(h)
    pred = acts[-1]
    loss = (-(target * pred.log())).sumAxis(1).sumAxis(0) / float(batch)   # CrossEntropy, Neural.fs:42-43
    # backward (softmax + cross-entropy with one-hot targets: dL/dz = (pred - target) / batch)
    dz = (pred - target) / float(batch)
    for li in range(len(params) - 1, -1, -1):
        w, b = params[li]
        h_in = acts[li]
        dw = dz.T @ h_in                      # [out, in]
        db = dz.sumAxis(0)
        if li > 0:
            dh = dz @ w                       # [batch, in]
            dz = dh * (1.0 - h_in * h_in)     # tanh'
        w.FillSubtract(w, dw * step)          # pars - step * grad, in place
        b.FillSubtract(b, db * step)
    return loss, pred


def _train_step_fused(x: Tensor, target: Tensor, params, step: float):
    batch = x.Shape[0]
    inv_b = float(batch)
    acts = [x]
    h = x
    for li, (w, b) in enumerate(params):
        z = h @ w.T
        if li < len(params) - 1:
            h = Tensor.fused(lambda zz, bb: (zz + bb).tanh(), z, b)
        else:
            z = z + b
            c = z.maxAxis(1).padRight()
            y = Tensor.fused(lambda zz, cc: (zz - cc).exp(), z, c)
            h = y / y.sumAxis(1).padRight()
        acts.append(h)
    pred = acts[-1]
    loss = Tensor.fused(lambda t, p: -(t * p.log()), target, pred).sumAxis(1).sumAxis(0) / float(batch)
    dz = Tensor.fused(lambda p, t: (p - t) / inv_b, pred, target)
    for li in range(len(params) - 1, -1, -1):
        w, b = params[li]
        h_in = acts[li]
        dw = dz.T @ h_in
        db = dz.sumAxis(0)
        if li > 0:
            dh = dz @ w
            dz = Tensor.fused(lambda d, hh: d * (1.0 - hh * hh), dh, h_in)
        w.FillFused(lambda ww, g: ww - g * step, w, dw)
        b.FillFused(lambda bb, g: bb - g * step, b, db)
    return loss, pred


def flops_per_step(batch: int, sizes) -> float:
    f = 0.0
    for li, (n_in, n_out) in enumerate(zip(sizes[:-1], sizes[1:])):
        f += 2.0 * batch * n_in * n_out          # forward
        f += 2.0 * batch * n_in * n_out          # dW
        if li > 0:
            f += 2.0 * batch * n_in * n_out      # dX
    return f


def main():
    import torch
    from deepnet_b200 import CudaTensor
    sizes, batch = (784, 4096, 4096, 10), 8192
    dev = CudaTensor.dev()
    dev.Init(0)
    dev.SetStream(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(5)
    params = [(CudaTensor.ofNumpy(w), CudaTensor.ofNumpy(b)) for w, b in init_params(rng, sizes)]
    xn, tn = synthetic_batch(rng, batch, sizes[0], sizes[-1])
    x, t = CudaTensor.ofNumpy(xn), CudaTensor.ofNumpy(tn)
    losses = []
    for _ in range(3):
        loss, _ = train_step(x, t, params, 0.01)
        losses.append(float(loss.Value))
    torch.cuda.synchronize()
    l0 = dev.LaunchCount()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    n = 10
    for _ in range(n):
        loss, _ = train_step(x, t, params, 0.01)
    e.record()
    e.synchronize()
    ms = s.elapsed_time(e) / n
    losses.append(float(loss.Value))
    fl = flops_per_step(batch, sizes)
    print(f"C5 MLP 784-4096-4096-10 batch 8192: {ms:.3f} ms/step, {fl / ms / 1e9:.1f} TFLOP/s (GEMM flops "
          f"{fl / 1e12:.3f} TFLOP/step), {(dev.LaunchCount() - l0) // n} kernel launches/step, losses {losses}")
    for _ in range(2):  # warm-up: the first launch of a kernel loads its module (CUDA lazy loading)
        train_step(x, t, params, 0.01, fused=True)
    torch.cuda.synchronize()
    l0 = dev.LaunchCount()
    s.record()
    for _ in range(n):
        loss, _ = train_step(x, t, params, 0.01, fused=True)
    e.record()
    e.synchronize()
    ms = s.elapsed_time(e) / n
    print(f"C5 MLP with FusedElemwise: {ms:.3f} ms/step, {fl / ms / 1e9:.1f} TFLOP/s, "
          f"{(dev.LaunchCount() - l0) // n} kernel launches/step, loss {float(loss.Value)}")


if __name__ == "__main__":
    main()
