#!/usr/bin/env python
"""Static instruction mix of a generated kernel (no GPU needed): lowers a BASELINE workload, lets
NVRTC cross-compile it for sm_100a and counts SASS opcodes with cuobjdump.

    PYTHONPATH=. python tools/sass_mix.py laplacian512|jacobi3d|jacobi3d-fused|simple1024|conv5|bilateral [BOUNDARY]

Prints registers / spills, the opcode histogram of the whole kernel and of its largest loop (the
instructions between the farthest backward branch and its target), and the markers that matter on
this path: UTMALDG (TMA loads), SYNCS (mbarrier), LDS/STS widths, STG widths, ACQBULK/PREEXIT.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Shape(object):
    def __init__(self, shape, dtype=numpy.float32):
        self.shape, self.dtype, self.ndim = tuple(shape), numpy.dtype(dtype), len(shape)


def plan_of(name, boundary):
    import bench
    from stencil_code_b200.backend.cuda import codegen_temporal, planner
    fused = name.endswith("-fused")
    workload = bench.make_workload(name.replace("-fused", ""), boundary, "cuda")
    stencil = workload["stencil"]
    args = (_Shape(workload["shape"]),) + tuple(_Shape(t.shape, t.dtype) for t in bench.stencil_tables(stencil))
    spec = stencil.specializer
    plan = spec.transform(spec.args_to_subconfig(args))
    if fused:
        plan = codegen_temporal.make_plan(plan.program, plan.domain, planner.B200, plan.options)
    return plan


def sass_of(plan, extra_options=()):
    from stencil_code_b200.backend.cuda import runtime
    os.environ.setdefault("STENCIL_CUDA_CACHE", "off")
    module = runtime.compile_source(plan.source, "sass_mix.cu", list(plan.options) + list(extra_options))
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as handle:
        handle.write(module.cubin())
        path = handle.name
    try:
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
        usage = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    finally:
        os.unlink(path)
    return sass, usage


LINE = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")


def instructions(sass, function):
    """[(address, opcode, operands)] of one function"""
    out, active = [], False
    for line in sass.splitlines():
        if line.strip().startswith("Function :"):
            active = line.strip().split(":")[1].strip() == function
            continue
        if active:
            match = LINE.search(line)
            if match:
                out.append((int(match.group(1), 16), match.group(2), match.group(3)))
    return out


def largest_loop(code):
    best = None
    for address, opcode, operands in code:
        if opcode.startswith("BRA"):
            target = re.search(r"0x([0-9a-f]+)", operands)
            if target:
                target = int(target.group(1), 16)
                if target < address and (best is None or address - target > best[1] - best[0]):
                    best = (target, address)
    return best


def histogram(code):
    counts = collections.Counter()
    for _, opcode, _ in code:
        counts[opcode] += 1
    return counts


def show(title, counts, top=28):
    total = sum(counts.values())
    print("{} ({} instructions)".format(title, total))
    base = collections.Counter()
    for opcode, count in counts.items():
        base[opcode.split(".")[0]] += count
    print("  by base opcode: " + ", ".join("{} {}".format(k, v) for k, v in base.most_common(top)))
    wide = {k: v for k, v in counts.items() if k.split(".")[0] in ("LDS", "STS", "STG", "LDG", "UTMALDG", "SYNCS",
                                                                    "ACQBULK", "PREEXIT", "BAR", "FSEL", "SEL")}
    print("  memory / sync:  " + ", ".join("{} {}".format(k, v) for k, v in sorted(wide.items())))
    floating = sum(v for k, v in base.items() if k in ("FADD", "FMUL", "FFMA", "DADD", "DMUL", "DFMA"))
    print("  floating point: {} ({:.0f} %)".format(floating, 100.0 * floating / max(1, total)))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "laplacian512"
    boundary = sys.argv[2] if len(sys.argv) > 2 else None
    plan = plan_of(name, boundary)
    sass, usage = sass_of(plan)
    print("{}: strategy {}, function {}, config {}".format(name, plan.strategy, plan.function, (plan.notes or {}).get("config")))
    for line in usage.splitlines():
        if "REG:" in line:
            print("  " + line.strip())
    code = instructions(sass, plan.function)
    show("whole kernel", histogram(code))
    loop = largest_loop(code)
    if loop:
        body = [entry for entry in code if loop[0] <= entry[0] <= loop[1]]
        show("largest loop 0x{:x}..0x{:x}".format(*loop), histogram(body))


if __name__ == "__main__":
    main()
