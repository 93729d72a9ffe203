"""Consistency of the generated code paths at BASELINE.json's full sizes, device-resident.

(The comparison with the CPU oracle at these sizes is ``tests/test_gpu_baseline_oracle.py``; this
file adds what needs no host round trip.)  Two kinds of evidence:
  * the independently generated code paths -- 'direct' (one thread per point, global loads),
    'stream' (TMA pipeline) and, for iterated sweeps, the fused two-sweep kernel -- must agree
    bit for bit on the full grid (``sc_count_diff`` counts differing bit patterns on the device);
    'direct' itself is pinned to the oracle on small grids by ``test_gpu_parity.py``;
  * exact algebraic properties: a constant field c maps to (sum of weights) * c everywhere under
    'clamp', and scaling the input by 2 scales the output by exactly 2.
"""
import ctypes

import numpy
import pytest

pytestmark = pytest.mark.gpu


class Shape(object):
    def __init__(self, shape):
        self.shape, self.dtype = tuple(shape), numpy.dtype('float32')


class OnDevice(object):
    """helper: device buffers, fills, sweeps with a forced strategy, bitwise comparison"""

    def __init__(self):
        from stencil_code_b200.backend.cuda import runtime
        self.runtime = runtime
        self.device = runtime.Device.get(0)
        self.library = runtime.load_library()
        self.fill_geometry = type("L", (), dict(grid=(148 * 8, 1, 1), block=(256, 1, 1), cluster=(1, 1, 1), smem=0))()

    def uniform(self, points, seed, scale=1.0):
        buffer = self.device.alloc(points * 4)
        self.runtime.launch(self.device.utils(), "sc_fill_uniform", self.fill_geometry,
                            [ctypes.c_void_p(buffer.ptr), ctypes.c_uint64(points), ctypes.c_uint64(seed),
                             ctypes.c_uint64(0), ctypes.c_float(scale)], self.device.stream)
        return buffer

    def constant(self, points, value):
        buffer = self.device.alloc(points * 4)
        bits = int(numpy.float32(value).view(numpy.uint32))
        self.runtime.check(self.library.sc_memset32_async(ctypes.c_void_p(buffer.ptr), bits, points, self.device.stream))
        return buffer

    def differing(self, a, b, points):
        cell = self.device.alloc(256)
        self.runtime.check(self.library.sc_memset32_async(ctypes.c_void_p(cell.ptr), 0, 2, self.device.stream))
        self.runtime.launch(self.device.utils(), "sc_count_diff", self.fill_geometry,
                            [ctypes.c_void_p(a.ptr), ctypes.c_void_p(b.ptr), ctypes.c_uint64(points),
                             ctypes.c_void_p(cell.ptr)], self.device.stream)
        host = numpy.zeros(1, numpy.uint64)
        self.runtime.check(self.library.sc_memcpy_d2h_async(ctypes.c_void_p(host.ctypes.data), ctypes.c_void_p(cell.ptr),
                                                            8, self.device.stream))
        self.device.synchronize()
        cell.release()
        return int(host[0])

    def tables(self, stencil):
        arrays = ()
        if hasattr(stencil, "intensity_lut"):
            arrays = (stencil.distance_lut, stencil.intensity_lut)
        elif hasattr(stencil, "coefficients"):
            arrays = (stencil.coefficients,)
        buffers = []
        for array in arrays:
            array = numpy.ascontiguousarray(array)
            buffer = self.device.alloc(array.nbytes)
            self.runtime.check(self.library.sc_memcpy_h2d_async(ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(array.ctypes.data),
                                                                array.nbytes, self.device.stream))
            buffers.append(buffer)
        self.device.synchronize()
        return arrays, buffers

    def concrete(self, make, shape, strategy, monkeypatch):
        if strategy is None:
            monkeypatch.delenv("STENCIL_CUDA_STRATEGY", raising=False)
        else:
            monkeypatch.setenv("STENCIL_CUDA_STRATEGY", strategy)
        stencil = make()
        arrays, buffers = self.tables(stencil)
        return stencil.specializer.specialize((Shape(shape),) + tuple(arrays)), buffers


def head(device_helper, buffer, count=8):
    host = numpy.zeros(count, numpy.float32)
    device_helper.runtime.check(device_helper.library.sc_memcpy_d2h_async(
        ctypes.c_void_p(host.ctypes.data), ctypes.c_void_p(buffer.ptr), host.nbytes, device_helper.device.stream))
    device_helper.device.synchronize()
    return host


def make_conv5(mode):
    from stencil_code_b200.library.convolution import ConvolutionFilter
    kernel = numpy.ones((5, 5), numpy.int64)
    kernel[2, 2] = -4
    return ConvolutionFilter(convolution_array=kernel, boundary_handling=mode)


@pytest.mark.parametrize("name,shape,modes,scale", [
    ("laplacian", (512, 512, 512), ("clamp", "copy"), 1.0),                 # C2
    ("conv5", (16384, 16384), ("zero", "clamp", "wrap", "copy"), 1.0),      # C4
    ("bilateral", (8192, 8192), ("clamp",), 255.0),                         # C3
])
def test_stream_equals_direct_at_full_size(name, shape, modes, scale, monkeypatch):
    from stencil_code_b200.library.laplacian import LaplacianKernel
    from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter
    helper = OnDevice()
    points = int(numpy.prod(shape))
    grid = helper.uniform(points, seed=7, scale=scale)
    for mode in modes:
        make = {"laplacian": lambda: LaplacianKernel(boundary_handling=mode),
                "conv5": lambda: make_conv5(mode),
                "bilateral": lambda: BetterBilateralFilter(3, 70)}[name]
        results = []
        for strategy in ("direct", "stream"):
            concrete, tables = helper.concrete(make, shape, strategy, monkeypatch)
            assert concrete.plan.strategy == strategy
            out = helper.device.alloc(points * 4)
            concrete.launch([grid.ptr] + [t.ptr for t in tables], out.ptr)
            results.append(out)
        assert helper.differing(results[0], results[1], points) == 0, (name, mode)
        for buffer in results:
            buffer.release()
    grid.release()


def test_jacobi3d_full_grid_paths_agree(monkeypatch):
    """C5: 2048x1024x1024 (2^31 points, 64-bit indexing).  direct == stream for one sweep, and
    two single sweeps == one fused double sweep."""
    from stencil_code_b200.library.jacobi_3d import Jacobi3D
    helper = OnDevice()
    shape = (2048, 1024, 1024)
    points = int(numpy.prod(shape))
    grid = helper.uniform(points, seed=11)
    a, b, c = (helper.device.alloc(points * 4) for _ in range(3))
    direct, _ = helper.concrete(lambda: Jacobi3D(1.0, 0.25), shape, "direct", monkeypatch)
    stream, _ = helper.concrete(lambda: Jacobi3D(1.0, 0.25), shape, "stream", monkeypatch)
    direct.launch([grid.ptr], a.ptr)
    stream.launch([grid.ptr], b.ptr)
    assert helper.differing(a, b, points) == 0
    stream.launch([b.ptr], a.ptr)                        # a = two single sweeps
    fused = stream.temporal()
    assert fused is not None
    stream._launch_plan(fused[0], fused[1], [grid.ptr], c.ptr, None, None, helper.device.stream)
    assert helper.differing(a, c, points) == 0
    for buffer in (grid, a, b, c):
        buffer.release()
    helper.device.trim()


def test_exact_properties_at_full_size(monkeypatch):
    from stencil_code_b200.library.laplacian import LaplacianKernel
    helper = OnDevice()
    shape = (512, 512, 512)
    points = int(numpy.prod(shape))
    concrete, _ = helper.concrete(lambda: LaplacianKernel(), shape, None, monkeypatch)
    # a constant field under clamp: -4c + 6c = 2c at every point, exactly
    constant = helper.constant(points, 3.0)
    expected = helper.constant(points, 6.0)
    out = helper.device.alloc(points * 4)
    concrete.launch([constant.ptr], out.ptr)
    assert helper.differing(out, expected, points) == 0
    # scaling by a power of two commutes with the stencil bit for bit
    x = helper.uniform(points, seed=5, scale=1.0)
    x2 = helper.uniform(points, seed=5, scale=2.0)
    y, y2 = helper.device.alloc(points * 4), helper.device.alloc(points * 4)
    concrete.launch([x.ptr], y.ptr)
    concrete.launch([x2.ptr], y2.ptr)
    concrete_scale, _ = helper.concrete(lambda: Doubler(), shape, None, monkeypatch)
    doubled = helper.device.alloc(points * 4)
    concrete_scale.launch([y.ptr], doubled.ptr)
    assert helper.differing(doubled, y2, points) == 0
    for buffer in (constant, expected, out, x, x2, y, y2, doubled):
        buffer.release()


from stencil_code_b200.stencil_kernel import Stencil   # noqa: E402


class Doubler(Stencil):
    neighborhoods = [[(0, 0, 0)]]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            out_grid[x] = 2 * in_grid[x]
