"""deepnet_b200 — B200-native CUDA device backend for Deep.Net's Tensor library.

The product is libdeepnet_b200.so (C ABI in include/dn_tensor.h, kernels in deepnet_b200/csrc). This package is
the host side above that ABI: the reference's ITensorDevice / ITensorStorage / ITensorBackend boundary and the
thin slice of the Tensor<'T> frontend that feeds it. Importing the package does not load the library; touching
`CudaTensor.dev()` does, and fails loudly if it has not been built. There is no CPU fallback.
"""
from . import dtypes, layout
from .backend import (ITensorDevice, ITensorStorage, NativeTensorBackend, TensorCudaBackend, TensorCudaDevice,
                      TensorCudaStorage, TensorStagingDevice)
from .layout import NotFound, TensorLayout
from .native import CudaException, NotSupportedException, OutOfCudaMemoryException, SingularMatrixException
from .tensor import CudaTensor, NoMask, Tensor

__all__ = ["dtypes", "layout", "Tensor", "CudaTensor", "NoMask", "NotFound", "TensorLayout", "ITensorDevice",
           "ITensorStorage", "NativeTensorBackend", "TensorCudaBackend", "TensorCudaDevice", "TensorCudaStorage",
           "TensorStagingDevice", "CudaException", "NotSupportedException", "OutOfCudaMemoryException", "SingularMatrixException"]
