"""3-d Jacobi sweep ``out = alpha * in + beta * sum(6 face neighbours)``.

The intent of reference ``examples/jacobi_3d.py:11-22`` (which is written against a removed API)
restated on the current one, as SURVEY.md section 8 row K5 prescribes: alpha and beta reach the
generated code as double literals (``self.alpha`` / ``self.beta`` fold exactly like entries of
the ``constants`` property, reference ``backend/stencil_backend.py:127-130``).
"""
from ..neighborhood import Neighborhood
from ..stencil_kernel import Stencil


class Jacobi3D(Stencil):
    neighborhoods = [Neighborhood.von_neuman_neighborhood(radius=1, dim=3, include_origin=False)]

    def __init__(self, alpha=1.0, beta=0.25, backend='cuda', boundary_handling='clamp'):
        self.alpha = alpha
        self.beta = beta
        super(Jacobi3D, self).__init__(backend=backend, boundary_handling=boundary_handling)

    @property
    def constants(self):
        return {'alpha': self.alpha, 'beta': self.beta}

    def kernel(self, in_grid, out_grid):
        for x in self.interior_points(out_grid):
            out_grid[x] = self.alpha * in_grid[x]
            for y in self.neighbors(x, 0):
                out_grid[x] += self.beta * in_grid[y]
