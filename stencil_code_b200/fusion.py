"""Composition of stencils: ``fuse(a, b, ...)(grid)`` computes ``...(b(a(grid)))``.

The reference explores this as "fusing two successive calls to a stencil specializer ... to reduce
the reads and writes to global memory by leveraging temporal locality", recomputing the ghost zone
of the intermediate grid per block (``ipython_notebooks/fusing_stencils/Fusing Stencils.rst``;
hooks: ``composable = True`` ``stencil_kernel.py:463``, ``get_ir_nodes`` ``:421-446``).  Here:

* the grid stays on the device between the stages (one host->device and one device->host copy for
  numpy input, none for torch CUDA tensors);
* two consecutive stages whose stencils are "stars in z" (off-centre taps only in the middle
  plane, rank 3, float32, the grid their only argument) run as ONE launch of the fused two-sweep
  kernel (``backend/cuda/codegen_temporal.py``), the first stencil producing the intermediate plane
  in shared memory / registers and the second consuming it: one read and one write of HBM for both;
* every other stage is one ordinary launch.

The result is bit-identical to calling the stencils one after the other.
"""
import ctypes

import numpy

from .stencil_exception import StencilException


class FusedStencils(object):
    def __init__(self, *stencils):
        if len(stencils) < 1:
            raise StencilException("fuse() needs at least one stencil")
        backends = {stencil.backend_name for stencil in stencils}
        if len(backends) != 1 or not backends <= {'cuda', 'python'}:
            raise StencilException("fuse() takes stencils of one backend, 'cuda' or 'python'; got {}".format(
                sorted(backends)))
        self.stencils = tuple(stencils)
        self.backend_name = backends.pop()
        self.last_schedule = None          # e.g. ['pair', 'single'] after a call (for tests / curiosity)

    def __call__(self, grid, tables=None):
        """``tables``: per stage, the extra array arguments of that stencil (default: none)."""
        tables = [tuple(t) for t in tables] if tables is not None else [()] * len(self.stencils)
        if len(tables) != len(self.stencils):
            raise StencilException("tables must hold one tuple per fused stencil")
        if self.backend_name == 'python':
            for stencil, extra in zip(self.stencils, tables):
                grid = stencil(grid, *extra)
            return grid
        from .backend.cuda import runtime
        stages = [stencil.specializer.specialize((grid,) + extra) for stencil, extra in zip(self.stencils, tables)]
        for concrete, extra in zip(stages, tables):
            concrete._check_args((grid,) + extra)
        if runtime._is_torch_tensor(grid):
            stages[0]._require_device_tensor(grid)
            return self._run_torch(stages, grid, tables)
        return self._run_numpy(stages, grid, tables)

    # ---- scheduling ----------------------------------------------------------------------------
    def _schedule(self, stages, tables):
        steps, k = [], 0
        while k < len(stages):
            pair = None
            if k + 1 < len(stages) and not tables[k] and not tables[k + 1]:
                pair = stages[k].pair_with(stages[k + 1])
            if pair is not None:
                steps.append(('pair', k, pair))
                k += 2
            else:
                steps.append(('single', k, None))
                k += 1
        self.last_schedule = [kind for kind, _, _ in steps]
        return steps

    def _execute(self, stages, steps, source, buffers, table_pointers, stream):
        """Runs the steps; ``source`` (the caller's grid) is only read.  Returns the pointer that
        holds the result (one of ``buffers``)."""
        current = source
        for index, (kind, k, pair) in enumerate(steps):
            target = buffers[index % 2]
            if kind == 'pair':
                stages[k]._launch_plan(pair[0], pair[1], [current], target, None, None, stream)
            else:
                stages[k].launch([current] + table_pointers[k], target, stream=stream)
            current = target
        return current

    def _run_torch(self, stages, grid, tables):
        import torch
        stream = ctypes.c_void_p(torch.cuda.current_stream(grid.device).cuda_stream)
        from .backend.cuda.runtime import _device_tensors
        source = grid.contiguous()
        table_tensors = [_device_tensors((source,) + extra)[1:] for extra in tables]
        steps = self._schedule(stages, tables)
        outputs = [torch.empty_like(source) for _ in range(min(2, len(steps)))]
        final = self._execute(stages, steps, source.data_ptr(), [t.data_ptr() for t in outputs],
                              [[t.data_ptr() for t in extra] for extra in table_tensors], stream)
        return outputs[0] if final == outputs[0].data_ptr() else outputs[1]

    def _run_numpy(self, stages, grid, tables):
        from .backend.cuda.runtime import check, load_library
        library = load_library()
        device = stages[0].device
        check(library.sc_set_device(device.index))
        stream = device.stream
        array = numpy.ascontiguousarray(grid)
        steps = self._schedule(stages, tables)
        held = []

        def upload(host):
            host = numpy.ascontiguousarray(host)
            buffer = device.alloc(host.nbytes)
            held.append((buffer, host))
            check(library.sc_memcpy_h2d_async(ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(host.ctypes.data),
                                              host.nbytes, stream))
            return buffer.ptr
        source = upload(array)
        table_pointers = [[upload(t) for t in extra] for extra in tables]
        outputs = [device.alloc(array.nbytes) for _ in range(min(2, len(steps)))]
        final = self._execute(stages, steps, source, [b.ptr for b in outputs], table_pointers, stream)
        result = device.pinned_empty(array.shape, array.dtype)
        check(library.sc_memcpy_d2h_async(ctypes.c_void_p(result.ctypes.data), ctypes.c_void_p(final),
                                          result.nbytes, stream))
        device.synchronize(stream)
        for buffer, _ in held:
            buffer.release()
        for buffer in outputs:
            buffer.release()
        return result


def fuse(*stencils):
    """``fuse(a, b, ...)`` -> callable with ``fused(grid) == ...(b(a(grid)))``, bit for bit."""
    return FusedStencils(*stencils)
