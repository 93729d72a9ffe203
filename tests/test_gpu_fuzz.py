"""Seeded fuzzing of the code generators: random stencils, both strategies, bit-exact vs the oracle.

Each case draws a rank (2 or 3), 1-3 neighborhoods of random (also one-sided) offsets, an optional
initialising statement, one accumulate statement per neighborhood with a random operator and a
random kind of weight (C int literal, power-of-two double, "odd" double such as 0.3 that must stay
in double, a table entry indexed by the Manhattan distance, a weight that multiplies a second input
grid), a boundary mode and an awkward shape.  The stencil class is written to a temporary module
(the front end reads its source), run through `direct` and `stream`, and compared with the CPU
oracle on the full grid.
"""
import importlib.util
import os
import random

import numpy
import pytest

from cases import oracle_output
from stencil_code_b200.stencil_exception import StencilException

pytestmark = pytest.mark.gpu

TEMPLATE = '''
from stencil_code_b200.stencil_kernel import Stencil


class Fuzzed(Stencil):
    neighborhoods = {neighborhoods!r}

    def distance(self, x, y):
        return sum([abs(x[i] - y[i]) for i in range(len(x))])

    def kernel(self, {params}):
        for x in self.interior_points(out_grid):
{body}
'''


def draw_case(seed):
    rng = random.Random(seed)
    dim = rng.choice((2, 3))
    radius = rng.choice((1, 1, 2, 3))
    two_grids = rng.random() < 0.3
    use_table = rng.random() < 0.4
    n_neighborhoods = rng.randint(1, 3)
    neighborhoods = []
    for _ in range(n_neighborhoods):
        count = rng.randint(1, 7)
        one_sided = rng.random() < 0.25
        offsets = []
        for _ in range(count):
            offset = tuple(rng.randint(0 if one_sided else -radius, radius) for _ in range(dim))
            offsets.append(offset)                      # duplicates are legal (README kernel)
        neighborhoods.append(offsets)
    lines = []
    init = rng.choice(("none", "scaled", "const", "copy"))
    if init == "scaled":
        lines.append("out_grid[x] = {} * in_grid[x]".format(rng.choice(("-4", "2", "0.5", "0.3", "1.0"))))
    elif init == "const":
        lines.append("out_grid[x] = {}".format(rng.choice(("1", "2.5"))))
    elif init == "copy":
        lines.append("out_grid[x] = in_grid[x]")
    for k in range(n_neighborhoods):
        op = rng.choice(("+=", "+=", "-="))
        source = "second_grid" if two_grids and rng.random() < 0.5 else "in_grid"
        kind = rng.choice(("int", "pow2", "odd", "table", "plain", "mixed"))
        if kind == "table" and not use_table:
            kind = "pow2"
        weight = {"int": rng.choice(("2", "3", "-1")), "pow2": rng.choice(("0.25", "2.0", "0.125")),
                  "odd": rng.choice(("0.3", "0.1", "1.7")), "table": "weights[self.distance(x, y)]"}.get(kind)
        if kind == "plain":
            expr = "{}[y]".format(source)
        elif kind == "mixed":
            expr = "({}[y] - in_grid[x]) * {}".format(source, rng.choice(("0.5", "3", "0.3")))
        elif rng.random() < 0.5:
            expr = "{} * {}[y]".format(weight, source)
        else:
            expr = "{}[y] * {}".format(source, weight)
        lines.append("for y in self.neighbors(x, {}):".format(k))
        lines.append("    out_grid[x] {} {}".format(op, expr))
    params = ["in_grid"] + (["second_grid"] if two_grids else []) + (["weights"] if use_table else []) + ["out_grid"]
    body = "\n".join(" " * 12 + line for line in lines)
    source = TEMPLATE.format(neighborhoods=neighborhoods, params=", ".join(params), body=body)
    if dim == 2:
        shape = (rng.choice((20, 37, 90)), rng.choice((64, 132, 260, 516)))
    else:
        shape = (rng.choice((8, 13, 20)), rng.choice((10, 33, 70)), rng.choice((64, 132, 260)))
    boundary = rng.choice(("clamp", "clamp", "zero", "copy"))
    table = numpy.array([1.0, 0.5, 0.25, -0.75, 0.3, 0.2, 0.1, 0.05, 0.025, 0.0125][:dim * radius + 1], numpy.float32)
    return dict(seed=seed, dim=dim, source=source, shape=shape, boundary=boundary, two_grids=two_grids,
                use_table=use_table, table=table)


def load_class(case, tmp_path):
    path = os.path.join(str(tmp_path), "fuzzed_{}.py".format(case["seed"]))
    with open(path, "w") as handle:
        handle.write(case["source"])
    spec = importlib.util.spec_from_file_location("fuzzed_{}".format(case["seed"]), path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module.Fuzzed


@pytest.mark.parametrize("seed", range(120))
def test_random_stencil_matches_oracle(seed, tmp_path, monkeypatch):
    case = draw_case(seed)
    Fuzzed = load_class(case, tmp_path)
    rng = numpy.random.RandomState(seed)
    args = [(rng.random_sample(case["shape"]) * 100).astype(numpy.float32)]
    if case["two_grids"]:
        args.append((rng.random_sample(case["shape"]) * 10).astype(numpy.float32))
    if case["use_table"]:
        args.append(case["table"])
    expected = oracle_output(Fuzzed(backend='python', boundary_handling=case["boundary"]), *args)
    ran = []
    for strategy in ("direct", "stream"):
        monkeypatch.setenv("STENCIL_CUDA_STRATEGY", strategy)
        try:
            actual = Fuzzed(backend='cuda', boundary_handling=case["boundary"])(*args)
        except StencilException as error:
            if strategy == "stream" and "cannot handle" in str(error):
                continue
            raise
        different = actual.view(numpy.uint32) != expected.view(numpy.uint32)
        assert not different.any(), "seed {} {} {}: {} elements differ, first at {}\n{}".format(
            seed, strategy, case["boundary"], int(different.sum()), tuple(numpy.argwhere(different)[0]), case["source"])
        ran.append(strategy)
    assert "direct" in ran


STAR_TEMPLATE = '''
from stencil_code_b200.stencil_kernel import Stencil


class Star(Stencil):
    neighborhoods = {neighborhoods!r}

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            out_grid[x] = {center} * in_grid[x]
            for y in self.neighbors(x, 0):
                out_grid[x] += {w0} * in_grid[y]
            for y in self.neighbors(x, 1):
                out_grid[x] {op}= in_grid[y] * {w1}
'''


@pytest.mark.parametrize("seed", range(16))
def test_random_star_stencil_through_the_fused_kernel(seed, tmp_path, monkeypatch):
    """Random stencils of the shape the fused two-sweep kernel accepts (off-centre taps only in
    the middle plane, centre taps at dz = -1, 0, +1): iterate(n) vs n oracle sweeps."""
    run_star_case(seed, (256, 388, 516), tmp_path, monkeypatch)


@pytest.mark.parametrize("seed", range(100, 112))
def test_random_star_stencil_on_rows_of_odd_length(seed, tmp_path, monkeypatch):
    """The same through iterate()'s padded-row buffers (rows that end inside a thread's quad)."""
    run_star_case(seed, (257, 390, 515), tmp_path, monkeypatch)


def run_star_case(seed, widths, tmp_path, monkeypatch):
    rng = random.Random(1000 + seed)
    in_plane = []
    for _ in range(rng.randint(2, 8)):
        in_plane.append((0, rng.randint(-2, 2), rng.randint(-3, 3)))
    neighborhoods = [[(1, 0, 0), (-1, 0, 0)], in_plane]
    source = STAR_TEMPLATE.format(neighborhoods=neighborhoods, center=rng.choice(("1.0", "-4", "0.5", "0.3")),
                                  w0=rng.choice(("0.25", "2", "0.1")), w1=rng.choice(("0.125", "3", "0.7")),
                                  op=rng.choice(("+", "-")))
    case = dict(seed=2000 + seed, source=source.replace("class Star", "class Fuzzed"))
    Star = load_class(case, tmp_path)
    shape = (rng.choice((9, 14, 23)), rng.choice((18, 40, 75)), rng.choice(widths))
    boundary = rng.choice(("clamp", "zero", "copy"))
    grid = numpy.random.RandomState(seed).random_sample(shape).astype(numpy.float32)
    monkeypatch.setenv("STENCIL_CUDA_TEMPORAL", "force")      # small planes: not "worthwhile"
    stencil = Star(backend='cuda', boundary_handling=boundary)
    if shape[2] % 4:
        monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)    # the grid itself: one thread per point
        concrete = stencil.specializer.specialize((grid,))._pitched_twin()
    else:
        monkeypatch.setenv("STENCIL_CUDA_STRATEGY", "stream")
        concrete = stencil.specializer.specialize((grid,))
    assert concrete is not None and concrete.temporal() is not None
    python_side = Star(backend='python', boundary_handling=boundary)
    expected = grid
    for sweeps in (1, 2, 3):
        expected = oracle_output(python_side, expected)
        if sweeps >= 2:
            actual = numpy.asarray(stencil.iterate(grid, sweeps))
            different = actual.view(numpy.uint32) != expected.view(numpy.uint32)
            assert not different.any(), "seed {} {} x{}: {} differ, first at {}\n{}".format(
                seed, boundary, sweeps, int(different.sum()), tuple(numpy.argwhere(different)[0]), source)


@pytest.mark.parametrize("seed", range(12))
def test_random_large_windows(seed, monkeypatch):
    """Convolutions with windows up to 15 x 17 taps: dense kernels take the tiled + re-rolled path
    (per-tap constant table), kernels with holes the unrolled shared-memory path, smaller ones
    the register window."""
    from stencil_code_b200.library.convolution import ConvolutionFilter
    rng = numpy.random.RandomState(500 + seed)
    ky, kx = [(13, 13), (15, 11), (9, 9), (7, 17), (11, 15), (5, 9)][seed % 6]
    kind = seed % 3
    if kind == 0:
        kernel = rng.randint(1, 6, size=(ky, kx)).astype(numpy.int64)             # dense ints
    elif kind == 1:
        kernel = numpy.round(rng.uniform(0.05, 2.0, size=(ky, kx)), 3)             # dense doubles
    else:
        kernel = numpy.round(rng.uniform(0.05, 2.0, size=(ky, kx)), 3)
        # holes (the tap set is no rectangle).  The pattern must be point-symmetric: like the
        # reference, ConvolutionFilter.distance looks the coefficient up at the MIRRORED offset
        holes = rng.random_sample((ky, kx)) < 0.3
        holes = holes | holes[::-1, ::-1]
        kernel[holes] = 0
        kernel[ky // 2, kx // 2] = 1.5
    boundary = ("clamp", "copy", "zero")[seed % 3]
    shape = [(40, 132), (70, 260), (33, 516)][seed % 3]
    grid = (rng.random_sample(shape) * 10).astype(numpy.float32)
    expected = oracle_output(ConvolutionFilter(kernel, backend='python', boundary_handling=boundary), grid)
    for strategy in ("direct", "stream"):
        monkeypatch.setenv("STENCIL_CUDA_STRATEGY", strategy)
        actual = numpy.asarray(ConvolutionFilter(kernel, boundary_handling=boundary)(grid))
        different = actual.view(numpy.uint32) != expected.view(numpy.uint32)
        assert not different.any(), "seed {} {} {}x{} {}: {} differ, first at {}".format(
            seed, strategy, ky, kx, boundary, int(different.sum()), tuple(numpy.argwhere(different)[0]))


@pytest.mark.parametrize("seed", range(200, 248))
def test_random_stencil_on_rows_of_odd_length(seed, tmp_path, monkeypatch):
    """The random stencils of the first test on grids whose rows are no multiple of 16 bytes:
    iterate() runs them through the streaming kernel on padded-row buffers."""
    case = draw_case(seed)
    if case["two_grids"]:
        pytest.skip("padded rows are for stencils with one grid argument")
    monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
    Fuzzed = load_class(case, tmp_path)
    shape = case["shape"][:-1] + (case["shape"][-1] + 1 + seed % 3,)
    rng = numpy.random.RandomState(seed)
    grid = rng.random_sample(shape).astype(numpy.float32)
    tables = (case["table"],) if case["use_table"] else ()
    python_side = Fuzzed(backend='python', boundary_handling=case["boundary"])
    once = oracle_output(python_side, grid, *tables)
    twice = oracle_output(python_side, once, *tables)
    stencil = Fuzzed(backend='cuda', boundary_handling=case["boundary"])
    twin = stencil.specializer.specialize((grid,) + tables)._pitched_twin()
    if twin is None:
        pytest.skip("the streaming kernel does not take this stencil")
    for sweeps, expected in ((1, once), (2, twice)):
        actual = stencil.iterate(grid, sweeps, *tables)
        different = actual.view(numpy.uint32) != expected.view(numpy.uint32)
        assert not different.any(), "seed {} {} x{}: {} differ, first at {}\n{}".format(
            seed, case["boundary"], sweeps, int(different.sum()), tuple(numpy.argwhere(different)[0]), case["source"])
