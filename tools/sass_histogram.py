"""Opcode histograms of the shipped SASS (cuobjdump on the object files the library is linked from): the evidence that
the GEMM runs on tcgen05 (UTCHMMA / UTMALDG / LDTM / UTCBAR), that the element-wise kernels use 256-bit accesses
(LDG.E.ENL2.256 / STG.E.ENL2.256) and what the reductions / compaction / scatter kernels issue.
Usage: python tools/sass_histogram.py [round-tag]  ->  profiles/sass_<tag>_*.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
TARGETS = [
    ("gemm.o", r"gemm_tf32_kernelILi256ELb0ELb0ELb0", "gemm_tf32_bn256"),
    ("gemm.o", r"gemm_tf32_kernelILi128ELb0ELb0ELb1", "gemm_3xtf32_bn128"),
    ("gemm.o", r"gemm_tf32_2cta_kernelILb0ELb0ELb0", "gemm_2cta"),
    ("ew_binary.o", r"ew_kernelINS_7BinaryFIfLi0EEELi8", "ew_add_f32_vec8"),
    ("ew_binary.o", r"ew_xpose_kernelINS_7BinaryFIfLi0", "ew_xpose_add_f32"),
    ("reduce_arg.o", r"reduce_rows_kernelINS_5ArgOpIfLb1", "reduce_rows_argmax_f32"),
    ("shard.o", r"reduce_rows_kernelINS_11MinMaxArgOpIfLb1", "reduce_rows_max_argmax_f32"),
    ("reduce.o", r"reduce_rows_kernelINS_5SumOpIaE", "reduce_rows_sum_i8"),
    ("reduce.o", r"reduce_rows_kernelINS_11MinMaxIntOpIaLb1", "reduce_rows_max_i8"),
    ("gemm.o", r"gemm_tf32_2cta_kernelILb0ELb0ELb1", "gemm_3xtf32_2cta"),
    ("shard.o", r"peer_push_kernel|peer_barrier_kernel", "shard_push_barrier"),
    ("index.o", r"compact_kernelINS_11CoordSink2DELb1ELi4", "compact_trueidx2d"),
    ("index.o", r"scatter_(hist|partition|accumulate)_kernel", "scatter_binned"),
    ("index.o", r"gather_kernelImLi1ELi1", "gather_u64_1d"),
]
line_re = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.*?);")
for obj, pat, name in TARGETS:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "build", "obj", obj)], capture_output=True, text=True).stdout
    on, funcs, hist = False, [], collections.Counter()
    for ln in sass.splitlines():
        if "Function :" in ln:
            on = re.search(pat, ln) is not None
            if on:
                funcs.append(ln.split("Function :")[1].strip())
            continue
        if on:
            m = line_re.match(ln)
            if m:
                toks = m.group(1).split()
                op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
                hist[op] += 1
    out = os.path.join(ROOT, "profiles", f"sass_{TAG}_{name}.txt")
    with open(out, "w") as fh:
        fh.write(f"# cuobjdump -sass build/obj/{obj}, functions matching /{pat}/ ({len(funcs)}):\n")
        for f in funcs[:8]:
            fh.write(f"#   {f}\n")
        for op, n in hist.most_common():
            fh.write(f"{n:8d}  {op}\n")
    key = {k: v for k, v in hist.items() if re.search(r"UTC|UTMA|LDTM|\.256|ATOMS|ATOMG|RED|SYNCS|MATCH|REDUX", k)}
    print(f"{os.path.basename(out)}: {len(funcs)} functions, {sum(hist.values())} instructions; {key}")
