"""The README's "simple kernel": sum of a 3x3 neighborhood (reference ``README.md:46-60``).

``SimpleKernel`` uses the Moore neighborhood the README evidently means; ``ReadmeSimpleKernel``
keeps the README's literal offset list, which repeats (-1, 0) and (-1, 1) and omits (0, -1) and
(1, -1) -- a regression case for duplicate offsets.
"""
from ..neighborhood import Neighborhood
from ..stencil_kernel import Stencil


class SimpleKernel(Stencil):
    neighborhoods = [Neighborhood.moore_neighborhood(radius=1, dim=2)]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            for y in self.neighbors(x, 0):
                out_grid[x] += in_grid[y]


class ReadmeSimpleKernel(SimpleKernel):
    neighborhoods = [[
        (-1, 1), (0, 1), (1, 1),
        (-1, 0), (0, 0), (1, 0),
        (-1, -1), (-1, 0), (-1, 1)
    ]]
