"""2-d "Jacobi" diagnostic with four one-point neighborhoods and integer weights
(reference ``library/jacobi_stencil.py:17-29``)."""
from ..stencil_kernel import Stencil


class Jacobi(Stencil):
    neighborhoods = [[(0, -1)], [(-1, 0)], [(0, 1)], [(1, 0)]]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            out_grid[x] = 1
            for y in self.neighbors(x, 0):
                out_grid[x] += 2 * in_grid[y]
            for y in self.neighbors(x, 1):
                out_grid[x] += 4 * in_grid[y]
            for y in self.neighbors(x, 2):
                out_grid[x] += 8 * in_grid[y]
            for y in self.neighbors(x, 3):
                out_grid[x] += 16 * in_grid[y]
