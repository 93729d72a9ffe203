"""3-d Laplacian benchmark (reference ``benchmarks/laplacian.py``: 16 x {32,64,128} x {1024,2048})."""
import numpy

from benchmarks.stencil_benchmarker import StencilBenchmarker, StencilTest, PrimedStencilTest
from stencil_code_b200.library.laplacian import LaplacianKernel


class LaplacianBenchmarker(StencilBenchmarker):
    def matrix_iterator(self):
        for height in (16,):
            for width in (32, 64, 128):
                for depth in [2 ** power for power in range(self.min_power, self.max_power)]:
                    yield numpy.random.random([height, width, depth]).astype(numpy.float32)


def main(backend='cuda', iterations=3):
    return LaplacianBenchmarker([
        StencilTest("unprimed-" + backend, stencil=LaplacianKernel, backend=backend),
        PrimedStencilTest("primed-" + backend, stencil=LaplacianKernel, backend=backend),
    ], iterations=iterations).run()


if __name__ == '__main__':
    main()
