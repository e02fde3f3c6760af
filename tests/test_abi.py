"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every symbol that
include/dn_tensor.h declares, the ctypes struct matches the C layout, and the product path fails loudly (no CPU
fallback) when there is no device."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from deepnet_b200 import CudaTensor, Tensor, TensorStagingDevice, dtypes, native
from deepnet_b200.native import NotSupportedException

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dn_tensor.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    api = native.product()
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(api.lib, s)]
    assert not missing, f"declared in dn_tensor.h but not exported: {missing}"
    assert set(native.ALL_PRODUCT_SYMBOLS) <= set(syms) | {"dn_last_error", "dn_launch_count", "dn_version"}


def test_struct_layout_matches_c(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dn_tensor.h"\nint main(){printf("%zu %zu %zu %zu '
                   '%zu %zu %zu\\n", sizeof(dn_tensor), offsetof(dn_tensor, base), offsetof(dn_tensor, offset), '
                   'offsetof(dn_tensor, ndims), offsetof(dn_tensor, dtype), offsetof(dn_tensor, shape), '
                   'offsetof(dn_tensor, stride));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    T = native.dn_tensor
    want = [C.sizeof(T), T.base.offset, T.offset.offset, T.ndims.offset, T.dtype.offset, T.shape.offset,
            T.stride.offset]
    assert got == want == [152, 0, 8, 16, 20, 24, 88]


def test_version_and_launch_counter():
    api = native.product()
    assert b"sm_100a" in api.lib.dn_version()
    assert api.lib.dn_launch_count() >= 0


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_fails_loudly_without_device():
    """No CUDA device: CudaInit.check raises (CudaBackend.fs:28-38) — nothing silently falls back to the CPU."""
    dev = CudaTensor.dev()
    with pytest.raises(native.CudaException):
        dev.Init(0)
    with pytest.raises((native.CudaException, MemoryError)):
        Tensor.zeros((4,), dtypes.DN_F32, dev)


def test_staging_tensors_have_no_compute_backend():
    t = Tensor.ofNumpy(np.arange(6, dtype=np.float32).reshape(2, 3))
    assert t.Dev == TensorStagingDevice.Instance()
    with pytest.raises(NotSupportedException):
        t + t


def test_product_package_does_not_import_the_oracle():
    code = ("import sys, deepnet_b200, deepnet_b200.tensor, deepnet_b200.backend, deepnet_b200.native;"
            "assert not [m for m in sys.modules if m.startswith('oracle')], 'oracle imported'")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "deepnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".in")):
                text = open(os.path.join(dirpath, f)).read()
                assert "host_oracle" not in text and "oracle." not in text.replace("oracle/", ""), f


# ---------------------------------------------------------------------------------------------------------------
# The F# P/Invoke binding (fsharp/Tensor.B200/Native.fs) cannot be compiled in this image (no dotnet); what can be
# checked without a compiler is checked here: every extern names an exported symbol, every declared entry point has
# an extern, and each pair agrees in arity and in the ABI class (pointer / int32 / int64) of every argument and of
# the return value — the properties a wrong P/Invoke signature would break.
# ---------------------------------------------------------------------------------------------------------------
def _c_prototypes():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"(dn_status|const char \*|int64_t)\s*(dn_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        def cls(a):
            a = a.strip()
            if "*" in a:
                return "ptr"
            if a.startswith("int32_t"):
                return "i32"
            if a.startswith("int64_t"):
                return "i64"
            raise AssertionError(f"{name}: unclassified C parameter {a!r}")
        params = [] if args.strip() in ("", "void") else [cls(a) for a in args.split(",")]
        protos[name] = ({"dn_status": "i32", "const char *": "ptr", "int64_t": "i64"}[ret], params)
    return protos


def _fsharp_externs():
    text = open(os.path.join(ROOT, "fsharp", "Tensor.B200", "Native.fs")).read()
    text = re.sub(r"//.*", "", text)
    externs = {}
    for ret, name, args in re.findall(r"extern\s+(\w+)\s+(dn_[a-z0-9_]+)\s*\(([^)]*)\)", text):
        def cls(a):
            ty = a.strip().rsplit(" ", 1)[0].strip()
            if "&" in ty or "[]" in ty or ty == "nativeint":
                return "ptr"
            return {"int": "i32", "int32": "i32", "int64": "i64"}[ty]
        params = [] if not args.strip() else [cls(a) for a in args.split(",")]
        externs[name] = ({"DnStatus": "i32", "nativeint": "ptr", "int64": "i64"}[ret], params)
    return externs


def test_fsharp_externs_match_the_header_and_the_library():
    api = native.product()
    protos, externs = _c_prototypes(), _fsharp_externs()
    assert len(protos) >= 70 and set(protos) == set(declared_symbols())
    assert sorted(set(protos) - set(externs)) == [], "declared in dn_tensor.h but not bound in Native.fs"
    assert sorted(set(externs) - set(protos)) == [], "bound in Native.fs but not declared in dn_tensor.h"
    for name, sig in externs.items():
        assert hasattr(api.lib, name), f"{name} is bound in Native.fs but not exported"
        assert sig == protos[name], f"{name}: Native.fs {sig} != dn_tensor.h {protos[name]}"


def test_fsharp_backend_references_only_defined_modules():
    """B200Backend.fs called B200Transfer / B200Index, which existed nowhere."""
    src = "".join(open(os.path.join(ROOT, "fsharp", "Tensor.B200", f)).read()
                  for f in ("Native.fs", "B200Backend.fs", "B200Shard.fs"))
    src = re.sub(r"//.*", "", src)
    defined = set(re.findall(r"^module (?:internal |private )?(\w+)", src, flags=re.M))
    used = set(re.findall(r"\b(B200\w+|Marshalling|Native)\.\w+", src))
    assert used <= defined, f"undefined F# modules referenced: {sorted(used - defined)}"
    natives = set(re.findall(r"Native\.(dn_[a-z0-9_]+)", src))
    assert natives <= set(_fsharp_externs()), sorted(natives - set(_fsharp_externs()))


def test_product_package_does_not_import_torch():
    """north_star: PyTorch is harness plumbing (tests, bench), not part of the product package."""
    code = ("import sys, deepnet_b200, deepnet_b200.shard, deepnet_b200.fused;"
            "assert 'torch' not in sys.modules, 'torch imported by the product package'")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
