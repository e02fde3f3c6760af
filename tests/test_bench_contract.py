"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the HostTensor restatement timed
on the host cores) prints ONE JSON line with the keys the driver reads, and the case list / byte accounting of the
headline workload is what DESIGN.md §5 states (33 calls per step, algorithmic bytes per SURVEY.md §8d)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-side", "256",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "GB/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_case_list_and_byte_accounting():
    sys.path.insert(0, ROOT)
    import bench
    from deepnet_b200 import Tensor, dtypes
    from oracle.host_tensor import HostTensor
    side = 64
    total_calls = 0
    for dt, npdt, has_sin in bench.dtype_list():
        s = np.dtype(npdt).itemsize
        a, b = (HostTensor.ofNumpy(np.ones((side, side), npdt)) for _ in range(2))
        row, col = HostTensor.ofNumpy(np.ones((1, side), npdt)), HostTensor.ofNumpy(np.ones((side, 1), npdt))
        c = Tensor.empty((side, side), dt, HostTensor.Dev)
        mask = Tensor.empty((side, side), dtypes.DN_BOOL, HostTensor.Dev)
        cases = bench.build_cases(Tensor, dt, a, b, c, row, col, mask, has_sin)
        total_calls += len(cases)
        by_name = {name.split(" ", 1)[1]: nb for name, nb, _, _ in cases}
        n = side * side
        assert by_name["add contiguous"] == 3 * n * s                      # binary: 3 s N
        assert by_name["add a + row[1,C]"] == (2 * n + side) * s           # broadcast operand: its own size
        assert by_name["copy a.T"] == 2 * n * s                            # unary / copy: 2 s N
        assert by_name["less a < b.T -> bool"] == (2 * s + 1) * n          # compare: 2 s N + N
        assert by_name["ifThenElse(mask, a, b)"] == (3 * s + 1) * n        # select: N + 3 s N
        assert by_name["add a[1:,1:] + b[1:,1:]"] == 3 * (side - 1) * (side - 1) * s
        for _, _, fn, tag in cases:                                         # every case runs on the host device too
            fn()
            assert isinstance(tag, str) and tag
    assert total_calls == 33
