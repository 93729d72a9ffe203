"""2-d heat flow over a stack of time planes: asymmetric neighborhoods, a repeated offset and a
``-=`` update (reference ``library/two_d_heat.py:20-32``)."""
from ..stencil_kernel import Stencil


class TwoDHeatFlow(Stencil):
    neighborhoods = [
        [(-1, 1, 0), (-1, -1, 0), (-1, 0, 1), (-1, 0, -1)],
        [(-1, 0, 0), (-1, 0, 0)],
    ]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            out_grid[x] = in_grid[x]
            for y in self.neighbors(x, 0):
                out_grid[x] += 0.125 * in_grid[y]
            for z in self.neighbors(x, 1):
                out_grid[x] -= 0.25 * in_grid[z]
