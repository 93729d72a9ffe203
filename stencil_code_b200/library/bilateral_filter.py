"""Bilateral filter with explicit look-up-table arguments
(reference ``library/bilateral_filter.py:9-33``).  No normalisation by the weight sum, as in the
reference; inputs must lie in [0, 255] so the intensity index stays inside ``filter_s``."""
import math

import numpy

from ..neighborhood import Neighborhood
from ..stencil_kernel import Stencil


class BilateralFilter(Stencil):
    def __init__(self, radius=3, backend='cuda'):
        super(BilateralFilter, self).__init__(
            neighborhoods=[Neighborhood.moore_neighborhood(radius=radius, dim=2)],
            backend=backend, should_unroll=False)

    def distance(self, x, y):
        return math.sqrt(sum([(x[i] - y[i]) ** 2 for i in range(0, len(x))]))

    def kernel(self, in_img, filter_d, filter_s, out_img):
        for i in self.interior_points(out_img):
            for j in self.neighbors(i, 0):
                out_img[i] += in_img[j] * filter_d[int(self.distance(i, j))] * \
                    filter_s[abs(int(in_img[i] - in_img[j]))]


def gaussian(stdev, length):
    """float32 table of ``scale * exp(-x^2 / (2 stdev^2))`` for x = 0..length-1, evaluated in
    Python doubles and rounded once (reference ``:27-33``)."""
    result = numpy.zeros(length).astype(numpy.float32)
    scale = 1.0 / (stdev * math.sqrt(2.0 * math.pi))
    divisor = -1.0 / (2.0 * stdev * stdev)
    for x in range(length):
        result[x] = scale * math.exp(float(x) * float(x) * divisor)
    return result
