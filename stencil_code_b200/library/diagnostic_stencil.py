"""Direction-coded 2-d stencil with double-typed weights 2, 4, 8, 16; on a grid of ones every
fully-inside point evaluates to 30 (reference ``library/diagnostic_stencil.py:6-21``)."""
from ..stencil_kernel import Stencil


class DiagnosticStencil(Stencil):
    neighborhoods = [[(0, -1)], [(0, 1)], [(-1, 0)], [(1, 0)]]

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            for y in self.neighbors(x, 0):
                out_grid[x] += 2.0 * in_grid[y]
            for y in self.neighbors(x, 1):
                out_grid[x] += 4.0 * in_grid[y]
            for y in self.neighbors(x, 2):
                out_grid[x] += 8.0 * in_grid[y]
            for y in self.neighbors(x, 3):
                out_grid[x] += 16.0 * in_grid[y]
