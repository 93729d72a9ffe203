"""Bilateral filter that owns its look-up tables
(reference ``library/better_bilateral_filter.py:13-71``): radius = 3 * sigma_d, distance table
of 2 * radius entries, intensity table of 256 entries, both passed behind the image on call."""
import math

from ..neighborhood import Neighborhood
from ..stencil_kernel import Stencil
from .bilateral_filter import gaussian as _gaussian


class BetterBilateralFilter(Stencil):
    def __init__(self, sigma_d=3, sigma_i=70, backend='cuda'):
        self.sigma_d = sigma_d
        self.sigma_i = sigma_i
        self.radius = 3 * self.sigma_d
        self.distance_lut = BetterBilateralFilter.gaussian(self.sigma_d, self.radius * 2)
        self.intensity_lut = BetterBilateralFilter.gaussian(self.sigma_i, 256)
        super(BetterBilateralFilter, self).__init__(
            neighborhoods=[Neighborhood.moore_neighborhood(radius=self.radius, dim=2)],
            backend=backend, should_unroll=False)

    def __call__(self, *args, **kwargs):
        return super(BetterBilateralFilter, self).__call__(
            args[0], self.distance_lut, self.intensity_lut, **kwargs)

    gaussian = staticmethod(_gaussian)

    def distance(self, x, y):
        return math.sqrt(sum([(x[i] - y[i]) ** 2 for i in range(0, len(x))]))

    def kernel(self, in_img, filter_d, filter_s, out_img):
        for i in self.interior_points(out_img):
            for j in self.neighbors(i, 0):
                out_img[i] += in_img[j] * filter_d[int(self.distance(i, j))] * \
                    filter_s[abs(int(in_img[i] - in_img[j]))]
