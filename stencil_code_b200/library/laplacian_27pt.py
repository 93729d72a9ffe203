"""27-point 3-d Laplacian whose weights come from a 4-entry table indexed by the Manhattan
distance of each tap (reference ``library/laplacian_27pt.py:9-30``)."""
from ..neighborhood import Neighborhood
from ..stencil_kernel import Stencil


class SpecializedLaplacian27(Stencil):
    neighborhoods = [Neighborhood.moore_neighborhood(radius=1, dim=3, include_origin=True)]

    def distance(self, x, y):
        return sum([abs(x[i] - y[i]) for i in range(len(x))])

    def kernel(self, source_data, factors, output_data):
        for x in self.interior_points(output_data):
            for n in self.neighbors(x, 0):
                output_data[x] += factors[self.distance(x, n)] * source_data[n]
