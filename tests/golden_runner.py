"""Runs the reference's documented known answers (tests/golden/doc_kats.json) on a given device."""
import json
import math
import os

import numpy as np

from deepnet_b200 import NoMask, Tensor, dtypes

_NP = {"int32": np.int32, "int64": np.int64, "double": np.float64, "single": np.float32, "bool": np.bool_}
HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases():
    with open(os.path.join(HERE, "golden", "doc_kats.json")) as f:
        return json.load(f)["cases"]


def _arr(spec):
    def conv(x):
        if isinstance(x, list):
            return [conv(y) for y in x]
        if isinstance(x, str):
            return {"inf": math.inf, "-inf": -math.inf, "nan": math.nan}[x]
        return x
    return np.array(conv(spec["data"]), dtype=_NP[spec["dtype"]])


def run_case(case, make):
    """`make(np_array) -> Tensor` places data on the device under test. Returns a numpy array / scalar / tuple."""
    t = {k: make(_arr(v)) for k, v in case["in"].items()}
    dev = next(iter(t.values())).Dev if t else make(np.zeros(1)).Dev
    op = case["op"]
    ax = case.get("axis")
    if op == "convert": r = t["a"].convert(dtypes.from_numpy(_NP[case["to"]]))
    elif op in ("sgn", "round", "sin", "isFinite"): r = getattr(t["a"], op)()
    elif op == "mod": r = t["a"] % t["b"]
    elif op == "pow": r = t["a"] ** t["b"]
    elif op == "add": r = t["a"] + t["b"]
    elif op == "mulScalar": r = t["a"] * case["value"]
    elif op == "eq": r = t["a"].eq(t["b"])
    elif op == "not": r = ~t["a"]
    elif op == "and": r = t["a"] & t["b"]
    elif op in ("maxElemwise", "minElemwise"): r = getattr(Tensor, op)(t["a"], t["b"])
    elif op == "ifThenElse": r = Tensor.ifThenElse(t["c"], t["a"], t["b"])
    elif op == "gather": r = Tensor.gather([t["i0"], t["i1"]], t["src"])
    elif op == "gather_none": r = Tensor.gather([None, t["j1"]], t["src"])
    elif op == "scatter": r = Tensor.scatter([t["i0"], t["i1"]], case["shape"], t["src"])
    elif op in ("countTrueAxis", "sumAxis", "productAxis", "minAxis", "maxAxis", "argMinAxis", "argMaxAxis",
                "allAxis", "anyAxis"): r = getattr(t["a"], op)(ax)
    elif op == "countTrue": return t["a"].countTrue()
    elif op == "trueIdx": r = t["a"].trueIdx()
    elif op == "argMax": return t["a"].argMax()
    elif op == "findAxis": r = t["a"].findAxis(case["value"], ax)
    elif op == "tryFind": return t["a"].tryFind(case["value"])
    elif op == "maskedGet": r = t["a"].M(t["m"])
    elif op == "maskedGetGt": r = t["a"].M(t["a"].gt(case["value"]))
    elif op == "maskedGet2": r = t["a"].M(t["m0"], t["m1"])
    elif op == "maskedGetNoMask": r = t["a"].M(t["m0"], NoMask)
    elif op == "maskedSet":
        t["a"].SetM([t["m"]], t["v"])
        r = t["a"]
    elif op == "counting": r = Tensor.counting(dev, case["n"])
    elif op == "arange": r = Tensor.arange(dev, *case["args"])
    elif op == "linspace": r = Tensor.linspace(dev, *case["args"])
    elif op == "fillMultiplyInPlace":
        f3 = Tensor.zeros(t["d"].Shape, t["d"].DataType, dev)
        f3.FillMultiply(t["d"], t["e"])
        f3.FillMultiply(f3, t["e"])
        r = f3
    elif op == "sum": return t["a"].sum()
    elif op == "all": return t["a"].all()
    elif op == "dot": r = t["a"] @ t["b"]
    elif op == "invert": r = Tensor.invert(t["a"])
    else: raise KeyError(op)
    return r.toNumpy()


def check_case(case, got):
    out = case["out"]
    if "scalar" in out:
        assert got == out["scalar"], f"{case['name']} ({case['ref']}): {got} != {out['scalar']}"
    elif "tuple" in out:
        assert tuple(got) == tuple(out["tuple"]), f"{case['name']} ({case['ref']}): {got}"
    else:
        want = _arr(out)
        assert got.shape == want.shape and got.dtype == want.dtype, f"{case['name']}: {got.shape} {got.dtype}"
        if "rtol" in out or "atol" in out:
            np.testing.assert_allclose(got, want, rtol=out.get("rtol", 0), atol=out.get("atol", 0),
                                       err_msg=f"{case['name']} ({case['ref']})")
        else:
            np.testing.assert_array_equal(got, want, err_msg=f"{case['name']} ({case['ref']})")
