"""5x5 convolution against ``scipy.ndimage.convolve`` -- the README benchmark of the reference
(``benchmarks/convolve.py``, ``README.md:26-34``), repaired: the reference script imports a module
that no longer exists (``:5``) and never resets its accumulators between sizes (``:33,92-99``).

Prints, per size, the reference's block: scipy average, specialised average with and without
specialisation ("compile") time, plus a correctness check of the interior against scipy.
"""
from __future__ import print_function

import numpy

from benchmarks.stencil_benchmarker import Timer
from stencil_code_b200.library.convolution import ConvolutionFilter

STENCIL = numpy.array([[1, 1, 1, 1, 1], [1, 1, 1, 1, 1], [1, 1, -4, 1, 1], [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]])


class Convolution(ConvolutionFilter):
    def __init__(self, backend):
        super(Convolution, self).__init__(STENCIL, backend=backend)


def main(backend='cuda', sizes=(1024, 2048, 4096, 8192), iterations=10, out=print):
    from scipy.ndimage import convolve
    results = []
    for width in sizes:
        image = numpy.random.random([width, width]).astype(numpy.float32)
        primed = Convolution(backend=backend)
        reference = convolve(image, STENCIL, mode='constant', cval=0.0)
        inside = primed.interior_points_slice()
        numpy.testing.assert_allclose(primed(image)[inside], reference[inside], rtol=1e-5, atol=1e-4)
        totals = [0.0, 0.0, 0.0]                         # reset per size (the reference forgets to)
        for _ in range(iterations):
            with Timer() as t0:
                convolve(image, STENCIL, mode='constant', cval=0.0)
            with Timer() as t1:
                Convolution(backend=backend)(image)
            with Timer() as t2:
                primed(image)
            totals = [a + b for a, b in zip(totals, (t0.interval, t1.interval, t2.interval))]
        averages = [t / iterations for t in totals]
        out("---------- Results for dim {0}x{0} ----------".format(width))
        out("Numpy convolve avg: {0}".format(averages[0]))
        out("Specialized {0} with compile time avg: {1}".format(backend, averages[1]))
        out("Specialized {0} time avg without compile {1}".format(backend, averages[2]))
        out("---------------------------------------------")
        results.append((width,) + tuple(averages))
    return results


if __name__ == '__main__':
    main()
