"""Neighborhood helpers: offset lists that define a stencil's taps.

Behavioural mirror of reference ``stencil_code/neighborhood.py:7-125``.  The ORDER of the
returned offsets is part of the contract: the backend unrolls one statement per offset in list
order (reference ``backend/stencil_backend.py:81-90``), so the order fixes the floating-point
summation order and therefore the bits of the result.  Offsets come out in row-major
lexicographic order, first coordinate slowest, each coordinate running -r..+r.
"""
import itertools

import numpy


class Neighborhood(object):
    """Static constructors for common neighborhoods (reference ``neighborhood.py:7``)."""

    @staticmethod
    def origin_of_dim(dim):
        """The all-zero point of a ``dim``-dimensional space (ref ``neighborhood.py:13-14``)."""
        return (0,) * dim

    @staticmethod
    def flatten(nested):
        """Flatten arbitrarily nested *lists* (tuples are leaves), dropping empty lists.

        Same result as the reference's in-place splice loop (``neighborhood.py:17-35``); the
        container type of the argument is preserved.
        """
        def leaves(item):
            if isinstance(item, list):
                for sub in item:
                    for leaf in leaves(sub):
                        yield leaf
            else:
                yield item

        return type(nested)(leaf for element in nested for leaf in leaves(element))

    @staticmethod
    def neighborhood_generator(r, d, l=()):
        """Nested-list cube of side 2r+1 in ``d`` dimensions whose leaves are coordinate tuples
        (ref ``neighborhood.py:38-48``).  ``flatten`` of it is the lexicographic point list."""
        prefix = list(l)
        if d == 0:
            return tuple(prefix)
        return [Neighborhood.neighborhood_generator(r, d - 1, prefix + [x])
                for x in range(-r, r + 1)]

    @staticmethod
    def _cube(radius, dim):
        return list(itertools.product(range(-radius, radius + 1), repeat=dim))

    @staticmethod
    def von_neuman_neighborhood(radius=1, dim=2, include_origin=True):
        """Points whose taxi-cab (L1) distance from the origin is <= radius
        (ref ``neighborhood.py:51-62``; the one-"n" spelling is the reference's)."""
        points = [p for p in Neighborhood._cube(radius, dim)
                  if sum(abs(c) for c in p) <= radius]
        if not include_origin:
            points.remove(Neighborhood.origin_of_dim(dim))
        return points

    # correctly spelled alias; additive
    von_neumann_neighborhood = von_neuman_neighborhood

    @staticmethod
    def moore_neighborhood(radius=1, dim=2, include_origin=True):
        """Points whose every coordinate is within radius (ref ``neighborhood.py:65-73``)."""
        points = Neighborhood._cube(radius, dim)
        if not include_origin:
            points.remove(Neighborhood.origin_of_dim(dim))
        return points

    @staticmethod
    def compute_halo(point_list):
        """Per-dimension (low, high) ghost extents of a point list (ref ``neighborhood.py:76-94``).

        Like the reference, both entries hold max |offset| (the halo is kept symmetric).
        Returns None for an empty list.
        """
        if len(point_list) == 0:
            return None
        dim = len(point_list[0])
        extent = [0] * dim
        for point in point_list:
            for d in range(dim):
                extent[d] = max(extent[d], abs(int(point[d])))
        return tuple((e, e) for e in extent)

    @staticmethod
    def compute_from_indices(matrix, mid_point=None):
        """Neighborhood + coefficient list from a coefficient matrix (ref ``neighborhood.py:97-125``).

        Every non-zero entry becomes a neighbor at ``index - mid_point`` (mid_point defaults to
        ``shape // 2``); entries are visited in C (row-major) order, which fixes the summation
        order.  Returns ``(neighbor_points, coefficients, halo)``; coefficients keep their numpy
        scalar type, which later decides whether they print as C int or double literals.
        """
        matrix = numpy.array(matrix)
        if mid_point is None:
            mid_point = [extent // 2 for extent in matrix.shape]
        neighbor_points = []
        coefficients = []
        for index in numpy.ndindex(*matrix.shape):
            value = matrix[index]
            if value != 0:
                neighbor_points.append(
                    tuple(int(index[d] - mid_point[d]) for d in range(matrix.ndim)))
                coefficients.append(value)
        return neighbor_points, coefficients, Neighborhood.compute_halo(neighbor_points)
