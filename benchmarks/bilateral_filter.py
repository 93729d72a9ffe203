"""Bilateral filter benchmark (reference ``benchmarks/bilateral_filter.py``).  Inputs are scaled to
[0, 255] -- the reference feeds [0, 1) images, which only ever index entry 0 of the intensity
table; the filter is specified for 8-bit intensities (``demo/bilateral_filter.py``)."""
import numpy

from benchmarks.stencil_benchmarker import StencilBenchmarker, StencilTest, PrimedStencilTest
from stencil_code_b200.library.better_bilateral_filter import BetterBilateralFilter


class ImageBenchmarker(StencilBenchmarker):
    def matrix_iterator(self):
        for matrix in super(ImageBenchmarker, self).matrix_iterator():
            yield matrix * numpy.float32(255)


def main(backend='cuda', iterations=3):
    return ImageBenchmarker([
        StencilTest("unprimed-" + backend, stencil=BetterBilateralFilter, backend=backend),
        PrimedStencilTest("primed-" + backend, stencil=BetterBilateralFilter, backend=backend),
    ], iterations=iterations).run()


if __name__ == '__main__':
    main()
