"""Iterated 3-d Jacobi relaxation: the intent of the reference's ``examples/jacobi_3d.py`` (written
against a ``StencilKernel`` / ``neighbor_points`` API the reference no longer has) on the current
``Stencil`` API.

    python examples/jacobi_3d.py [--backend cuda|python] [--shape 256 256 256] [--iterations 100]

``alpha`` and ``beta`` reach the generated code as literals through the ``constants`` property
(reference ``stencil_kernel.py:561-563``); the loop ``g = stencil(g)`` a reference user would write
is ``stencil.iterate(g, n)`` here: the grid stays on the device, pairs of sweeps run fused.
"""
from __future__ import print_function

import argparse
import os
import sys
import time

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from stencil_code_b200.library.jacobi_3d import Jacobi3D      # noqa: E402


def main(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("--backend", default="cuda", choices=("cuda", "python"))
    parser.add_argument("--shape", type=int, nargs=3, default=(256, 256, 256))
    parser.add_argument("--iterations", type=int, default=100)
    parser.add_argument("--boundary", default="clamp", choices=("clamp", "zero", "copy", "wrap", "periodic"))
    args = parser.parse_args(argv)
    stencil = Jacobi3D(alpha=0.4, beta=0.1, backend=args.backend, boundary_handling=args.boundary)
    print("j.alpha = {}, j.beta = {}".format(stencil.alpha, stencil.beta))
    grid = numpy.random.RandomState(1).random_sample(tuple(args.shape)).astype(numpy.float32)
    begin = time.time()
    result = stencil.iterate(grid, args.iterations)
    seconds = time.time() - begin
    points = float(numpy.prod(args.shape)) * args.iterations
    print("{} sweeps of {} on backend '{}': {:.3f} s ({:.3f} GStencil/s incl. compile and copies), "
          "mean {:.6f} -> {:.6f}".format(args.iterations, "x".join(map(str, args.shape)), args.backend, seconds,
                                         points / seconds / 1e9, float(grid.mean()), float(numpy.asarray(result).mean())))
    return result


if __name__ == "__main__":
    main()
