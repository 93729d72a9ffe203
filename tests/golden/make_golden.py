"""Generate tests/golden/*.npz from the UNMODIFIED reference, run in this container.

The reference's specialising back ends need ctree / pycl / hindemith (absent), but its
``backend='python'`` path (``stencil_code/stencil_kernel.py:545-559``) needs none of them at run
time -- they are only imported at module level.  This script registers inert stand-in modules for
those three packages (every attribute is an empty class, nothing of them is ever called), imports
``stencil_code`` from ``/root/reference`` and runs the reference's own ``Stencil`` subclasses from
``stencil_code/library`` on seeded inputs.  Outputs are stored with the inputs.

Two kinds of input per case:
  * "exact": small-integer-valued float grids.  Every intermediate of every library kernel is then
    exactly representable in float32, so the result does not depend on whether numpy evaluates
    ``2.0 * float32`` in float32 (numpy >= 2) or float64 (numpy 1.x, and C) -> bit-exact check.
  * "random": uniform random grids as the reference's tests use (x1000 or x255) -> checked to the
    reference's own tolerance (6 decimals, relative), because the Python path under numpy 2 rounds
    differently from C in the last ulp.

Run:  python tests/golden/make_golden.py      (needs /root/reference; not needed on the GPU box)
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
import warnings

import numpy

REFERENCE = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
MISSING = ("ctree", "pycl", "hindemith")


class _Inert(types.ModuleType):
    """Module whose every attribute is a fresh empty class (usable as a base class)."""
    __path__ = []
    __all__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        placeholder = type(name, (object,), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, placeholder)
        return placeholder


class _InertFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in MISSING:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Inert(spec.name)

    def exec_module(self, module):
        pass


def import_reference():
    sys.meta_path.insert(0, _InertFinder())
    sys.path.insert(0, REFERENCE)
    warnings.simplefilter("ignore", SyntaxWarning)
    for name in list(sys.modules):
        if name == "stencil_code" or name.startswith("stencil_code."):
            del sys.modules[name]
    import stencil_code.stencil_kernel as sk      # noqa: the reference's module
    assert sk.__file__.startswith(REFERENCE), sk.__file__
    return sk


def cases(sk):
    """(name, factory(boundary) -> reference stencil, extra-args, shape, scale, dims)."""
    from stencil_code.library.laplacian import LaplacianKernel
    from stencil_code.library.jacobi_stencil import Jacobi
    from stencil_code.library.diagnostic_stencil import DiagnosticStencil
    from stencil_code.library.two_d_heat import TwoDHeatFlow
    from stencil_code.library.laplacian_27pt import SpecializedLaplacian27
    from stencil_code.library.convolution import ConvolutionFilter
    from stencil_code.library.better_bilateral_filter import BetterBilateralFilter
    from stencil_code.library.bilateral_filter import BilateralFilter, gaussian
    from stencil_code.neighborhood import Neighborhood

    def plain(cls):
        return lambda boundary: cls(backend='python', boundary_handling=boundary)

    class Stencil1d(sk.Stencil):                  # test/test_boundary_handling.py:145-151
        neighborhoods = [[(-1,), (0,), (1,)]]

        def kernel(self, in_grid, out_grid):
            for x in self.interior_points(out_grid):
                for y in self.neighbors(x, 0):
                    out_grid[x] += 2 * in_grid[y]

    class Clamper(sk.Stencil):                    # test/test_boundary_handling.py:15-21
        neighborhoods = [Neighborhood.von_neuman_neighborhood(radius=1, dim=2)]

        def kernel(self, in_grid, out_grid):
            for p in self.interior_points(out_grid):
                for n in self.neighbors(p, 0):
                    out_grid[p] += in_grid[n]

    class ReadmeSimple(sk.Stencil):               # README.md:50-60
        neighborhoods = [[(-1, 1), (0, 1), (1, 1), (-1, 0), (0, 0), (1, 0),
                          (-1, -1), (-1, 0), (-1, 1)]]

        def kernel(self, in_grid, out_grid):
            for x in self.interior_points(out_grid):
                for y in self.neighbors(x, 0):
                    out_grid[x] += in_grid[y]

    conv5 = numpy.array([[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, -4, 1, 3],
                         [1, 1, 1, 1, 1], [1, 1, 1, 7, 1]])     # test/test_c_end_to_end.py:49-57
    coefficients27 = numpy.array([1.0, 0.5, 0.25, 0.125]).astype(numpy.float32)

    all_modes = ('clamp', 'zero', 'copy', 'wrap')
    return [
        ("laplacian", plain(LaplacianKernel), (), (9, 10, 12), 1000, all_modes),
        ("jacobi2d", plain(Jacobi), (), (17, 20), 1000, all_modes),
        ("diagnostic", plain(DiagnosticStencil), (), (16, 12), 1000, all_modes),
        ("two_d_heat", plain(TwoDHeatFlow), (), (6, 9, 8), 1000, all_modes),
        ("laplacian27", plain(SpecializedLaplacian27), (coefficients27,), (8, 9, 12), 255, all_modes),
        ("stencil1d", plain(Stencil1d), (), (24,), 1000, all_modes),
        ("clamper", plain(Clamper), (), (10, 10), 1000, all_modes),
        ("readme_simple", plain(ReadmeSimple), (), (12, 16), 1000, all_modes),
        # ConvolutionFilter hard-codes 'copy' (library/convolution.py:22-24)
        ("convolution5", lambda b: ConvolutionFilter(convolution_array=conv5, backend='python'),
         (), (20, 24), 1000, ('copy',)),
        # bilateral: clamp is the constructor default and cannot be changed
        ("better_bilateral", lambda b: BetterBilateralFilter(sigma_d=1, sigma_i=70, backend='python'),
         (), (20, 16), 255, ('clamp',)),
        ("bilateral_r3", lambda b: BilateralFilter(radius=3, backend='python'),
         (gaussian(1, 6), gaussian(70, 256)), (12, 16), 255, ('clamp',)),
    ]


def mallard(sk):
    """The reference's real-image fixture: ``demo/mallard.raw`` (800x800 8-bit grey) through the
    unmodified reference ``BetterBilateralFilter()`` -- the filter of ``demo/bilateral_filter.py:19``
    with its default sigmas 3 / 70 (radius 9) -- on its python backend.  A 64x64 crop keeps the
    fixture small (the pure-python backend needs ~10 s for it); the grid is filled as the demo
    does, ``in_grid[(x, y)] = pixels[y * width + x]`` (``demo/bilateral_filter.py:22-24``)."""
    from stencil_code.library.better_bilateral_filter import BetterBilateralFilter
    width = height = 800
    with open(os.path.join(REFERENCE, "demo", "mallard.raw"), "rb") as handle:
        pixels = numpy.frombuffer(handle.read(width * height), numpy.uint8)
    image = pixels.reshape(height, width).T.astype(numpy.float32)       # [x, y], as the demo indexes it
    x0, y0, n = 368, 300, 64                                           # a textured part of the bird
    crop = numpy.ascontiguousarray(image[x0:x0 + n, y0:y0 + n])
    out = BetterBilateralFilter(backend='python')(crop)
    assert out.dtype == numpy.float32 and out.shape == crop.shape
    path = os.path.join(HERE, "mallard_crop__clamp.npz")
    numpy.savez_compressed(path, image_in=crop, image_out=out, origin=numpy.array([x0, y0]),
                           mean_in=numpy.array(float(pixels.astype(numpy.float64).mean())))
    print("wrote", os.path.basename(path), crop.shape, "grey levels", int(crop.min()), "..", int(crop.max()))


def main():
    sk = import_reference()
    if "--mallard-only" in sys.argv:
        mallard(sk)
        return
    written = 0
    for name, factory, extra, shape, scale, modes in cases(sk):
        for mode in modes:
            stencil = factory(mode)
            rng = numpy.random.RandomState(sum(ord(c) for c in name + mode))
            exact_in = rng.randint(0, 8, size=shape).astype(numpy.float32)
            random_in = (rng.random_sample(shape) * scale).astype(numpy.float32)
            record = {"boundary": numpy.array(mode), "shape": numpy.array(shape)}
            for tag, grid in (("exact", exact_in), ("random", random_in)):
                args = (grid,) + tuple(extra)
                out = stencil(*args)
                assert out.dtype == numpy.float32 and out.shape == grid.shape
                record[tag + "_in"] = grid
                record[tag + "_out"] = out
            for i, table in enumerate(extra):
                record["extra{}".format(i)] = numpy.asarray(table)
            path = os.path.join(HERE, "{}__{}.npz".format(name, mode))
            numpy.savez_compressed(path, **record)
            written += 1
            print("wrote", os.path.basename(path), shape, mode)
    print(written, "fixtures")
    mallard(sk)


if __name__ == "__main__":
    main()
