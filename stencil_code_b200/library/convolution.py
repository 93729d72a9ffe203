"""Convolution with a user-supplied coefficient array (reference ``library/convolution.py:8-48``).

The neighborhood is the set of non-zero entries of ``convolution_array``; ``distance`` is
overridden to return the coefficient, which the backend folds to a literal per tap
(``backend/stencil_backend.py``).  As in the reference the coefficient applied to the neighbor at
``+offset`` is the array entry at ``-offset`` (a true convolution), the unused coefficient vector
is still passed as the second kernel argument, and the boundary handling defaults to 'copy'.
``boundary_handling`` is an additive keyword (the reference hard-codes 'copy', ``:22-24``).
"""
import numpy

from ..neighborhood import Neighborhood
from ..stencil_kernel import Stencil


class ConvolutionFilter(Stencil):
    def __init__(self, convolution_array=None, stride=1, backend='cuda',
                 boundary_handling='copy'):
        self.convolution_array = convolution_array
        neighbors, coefficients, _ = Neighborhood.compute_from_indices(convolution_array)
        self.neighbor_to_coefficient = dict(zip(neighbors, coefficients))
        self.coefficients = numpy.array(coefficients)
        self.stride = stride
        super(ConvolutionFilter, self).__init__(
            neighborhoods=[neighbors], backend=backend, boundary_handling=boundary_handling)

    def __call__(self, *args, **kwargs):
        if kwargs.get('iterations') is not None:
            return self.iterate(args[0], int(kwargs.pop('iterations')), **kwargs)
        return super(ConvolutionFilter, self).__call__(args[0], self.coefficients, **kwargs)

    def iterate(self, grid, iterations, **kwargs):
        """additive, like ``Stencil.iterate``: the coefficient vector is supplied as in ``__call__``"""
        return super(ConvolutionFilter, self).iterate(grid, iterations, self.coefficients, **kwargs)

    def distance(self, x, y):
        d = tuple(x[i] - y[i] for i in range(len(x)))
        return self.neighbor_to_coefficient[d]

    def kernel(self, input_grid, coefficients, output_grid):
        for point in self.interior_points(output_grid, stride=self.stride):
            for n in self.neighbors(point, 0):
                output_grid[point] += input_grid[n] * self.distance(point, n)
