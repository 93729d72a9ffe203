"""Compile and call the restated reference C (see ``c_emitter.py``).  Test infrastructure only.

``gcc -O2 -std=c99 -fPIC -shared [-fopenmp]``; shared objects are cached by source hash under
``oracle/_build/`` (git-ignored).  The call mirrors reference ``ConcreteStencil.__call__``
(``stencil_kernel.py:87-105``): a zero-initialised output of ``args[0]``'s shape and dtype is
appended together with ``byref(duration)``.
"""
import ctypes
import hashlib
import os
import subprocess
import time

import numpy

from .c_emitter import emit_c, OracleError

_BUILD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")
_LOADED = {}


def compile_c(source, openmp):
    tag = hashlib.sha1((source + ("omp" if openmp else "")).encode()).hexdigest()[:20]
    if tag in _LOADED:
        return _LOADED[tag]
    os.makedirs(_BUILD_DIR, exist_ok=True)
    so_path = os.path.join(_BUILD_DIR, "ref_{}.so".format(tag))
    if not os.path.exists(so_path):
        c_path = os.path.join(_BUILD_DIR, "ref_{}.c".format(tag))
        with open(c_path, "w") as handle:
            handle.write(source)
        tmp = so_path + ".tmp{}".format(os.getpid())
        command = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-w"]
        if openmp:
            command.append("-fopenmp")
        command += ["-o", tmp, c_path]
        done = subprocess.run(command, capture_output=True, text=True)
        if done.returncode != 0:
            raise OracleError("gcc failed:\n" + done.stderr[-2000:])
        os.replace(tmp, so_path)
    library = ctypes.CDLL(so_path)
    _LOADED[tag] = library
    return library


class ReferenceC(object):
    """One specialisation of the restated reference code, callable like the reference's
    concrete function."""

    def __init__(self, stencil, args, variant='c'):
        self.variant = variant
        self.shapes = [tuple(a.shape) for a in args]
        self.dtypes = [numpy.dtype(a.dtype) for a in args]
        self.source = emit_c(stencil, self.shapes, self.dtypes, variant)
        library = compile_c(self.source, openmp=variant in ('omp', 'c_parallel'))
        self.function = library.stencil_kernel
        self.function.restype = ctypes.c_int32
        self.function.argtypes = [ctypes.c_void_p] * (len(args) + 1) + [ctypes.POINTER(ctypes.c_float)]
        self.last_seconds = None

    def __call__(self, *args):
        arrays = [numpy.ascontiguousarray(a) for a in args]
        for array, shape, dtype in zip(arrays, self.shapes, self.dtypes):
            if tuple(array.shape) != shape or array.dtype != dtype:
                raise TypeError("argument does not match the specialised shape/dtype")
        output = numpy.zeros(self.shapes[0], self.dtypes[0])        # stencil_kernel.py:412-416
        duration = ctypes.c_float()
        pointers = [a.ctypes.data_as(ctypes.c_void_p) for a in arrays + [output]]
        start = time.perf_counter()
        self.function(*(pointers + [ctypes.byref(duration)]))
        self.last_seconds = time.perf_counter() - start
        return output


def reference_c(stencil, *args, **kwargs):
    """Output of the restated reference backend for ``stencil(*args)``.

    ``variant`` as in ``emit_c``.  Stencils that rewrite their argument list in ``__call__``
    (ConvolutionFilter, BetterBilateralFilter) are handled through ``reference_args``.
    """
    variant = kwargs.pop('variant', 'c')
    args = reference_args(stencil, args)
    return ReferenceC(stencil, args, variant)(*args)


def reference_args(stencil, args):
    """The argument tuple the stencil's own ``__call__`` override would pass on."""
    if hasattr(stencil, 'intensity_lut') and hasattr(stencil, 'distance_lut'):
        return (args[0], stencil.distance_lut, stencil.intensity_lut)
    if hasattr(stencil, 'neighbor_to_coefficient') and hasattr(stencil, 'coefficients'):
        return (args[0], stencil.coefficients)
    return tuple(args)
