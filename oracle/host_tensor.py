"""HostTensor device backed by the CPU oracle (oracle/host_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Gives the tests a `HostTensor.Dev` with the same frontend as `CudaTensor.Dev`, so a parity test reads like the
reference's own host-vs-CUDA comparisons (Tensor/Tensor.Test/CudaTests.fs:54-75, ML/AllTests/TestUtils.fs:28-47):
build the inputs on the host, run the same expression on both devices, compare.

Nothing in deepnet_b200/ imports this module; only tests/, __graft_entry__.smoke() and bench.py's CPU legs do.
"""
from __future__ import annotations

import os
import subprocess

import numpy as np

from deepnet_b200 import dtypes
from deepnet_b200 import layout as TL
from deepnet_b200.backend import ITensorDevice, ITensorStorage, NativeTensorBackend
from deepnet_b200.native import CApi
from deepnet_b200.tensor import Tensor

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libdn_oracle.so")
_API = None


def build() -> str:
    """Compile the oracle (g++, ~30 s) if the shared object is missing or stale."""
    src = os.path.join(_HERE, "host_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "dn_tensor.h")
    if (not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src)
            or os.path.getmtime(_LIB) < os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


def api() -> CApi:
    global _API
    if _API is None:
        _API = CApi(build(), "dno_", with_device_api=False)
        _API.lib.dno_set_threads_enabled.argtypes = [__import__("ctypes").c_int]
    return _API


def set_threads_enabled(enabled: bool) -> None:
    api().lib.dno_set_threads_enabled(1 if enabled else 0)


class TensorHostBackend(NativeTensorBackend):
    """TensorHostBackend<'T>, Tensor/Tensor/Host/HostBackend.fs:103-665 (restated in host_oracle.cpp)."""

    def GetItem(self, idx):
        return self.storage.array[TL.addr(idx, self.layout)]

    def SetItem(self, idx, value):
        self.storage.array[TL.addr(idx, self.layout)] = value

    def Transfer(self, trgt, src) -> bool:
        return False  # HostBackend.fs:210-211


class TensorHostStorage(ITensorStorage):
    """TensorHostStorage<'T>, HostBackend.fs:51-98: a flat managed array (numpy here)."""

    def __init__(self, array: np.ndarray, dev):
        self.Dev = dev
        self.array = array
        self.DataType = dtypes.from_numpy(array.dtype)
        self.DataSize = array.size

    def BasePtr(self) -> int:
        return self.array.ctypes.data

    def Backend(self, layout):
        return TensorHostBackend(layout, self, api())


class TensorHostDevice(ITensorDevice):
    """TensorHostDevice, HostBackend.fs:668-691 (`Id = "Host"`, `Zeroed = true`)."""
    Id = "Host"
    Zeroed = True
    _instance = None

    @classmethod
    def Instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def Create(self, nElems, dtype):
        if nElems > 2 ** 31 - 1:  # HostBackend.fs:54-58
            raise RuntimeError(f"Cannot create host tensor storage for {nElems} elements")
        return TensorHostStorage(np.zeros(max(1, int(nElems)), dtype=dtypes.to_numpy(dtype)), self)


class HostTensor:
    """module HostTensor, Tensor/Tensor/Host/HostFrontend.fs."""
    Dev = TensorHostDevice.Instance()

    @staticmethod
    def ofNumpy(arr: np.ndarray) -> Tensor:
        arr = np.ascontiguousarray(arr)
        flat = arr.reshape(-1).copy() if arr.size else np.zeros(1, arr.dtype)
        return Tensor(TL.newC(arr.shape), TensorHostStorage(flat, HostTensor.Dev))

    @staticmethod
    def zeros(shape, dtype) -> Tensor:
        return Tensor.zeros(shape, dtype, HostTensor.Dev)
