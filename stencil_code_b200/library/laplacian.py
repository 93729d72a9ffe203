"""7-point 3-d Laplacian (reference ``library/laplacian.py:11-22``; note the reference's centre
weight is -4 with six neighbours, kept as is)."""
from ..stencil_kernel import Stencil


class LaplacianKernel(Stencil):
    neighborhoods = [[(0, 0, 1), (0, 0, -1), (0, 1, 0), (0, -1, 0), (1, 0, 0), (-1, 0, 0)]]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(in_grid):
            out_grid[x] = -4 * in_grid[x]
            for y in self.neighbors(x, 0):
                out_grid[x] += in_grid[y]
