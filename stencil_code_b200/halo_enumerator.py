"""HaloEnumerator -- disjoint enumeration of the boundary shell of an n-d box.

Behavioural mirror of reference ``stencil_code/halo_enumerator.py:5-86``.  The shell is covered
dimension by dimension: for dimension k the low slab ``[0, h_k)`` and the high slab
``[n_k - h_k, n_k)`` are emitted with every dimension < k already restricted to its interior
(so no point is produced twice) and every dimension > k spanning its full extent.  It serves
``Stencil.halo_points`` (the python backend's 'copy' mode); the CUDA kernels handle their halo
points themselves and do not use it.
"""
import itertools


class HaloEnumerator(object):
    def __init__(self, halo, shape):
        """
        :param halo: ghost depth per dimension
        :param shape: extent per dimension of the box to enumerate
        """
        assert len(halo) == len(shape), \
            "HaloEnumerator halo({}) and shape({}) must be same size".format(halo, shape)
        self.halo = halo
        self.shape = shape

    def slabs(self):
        """Yield one tuple of per-dimension ``range`` objects per slab, in emission order.

        Two slabs (low, high) per dimension; this is the device-friendly form of the cover.
        """
        dims = len(self.shape)
        for k in range(dims):
            others_lo = [range(self.halo[d], self.shape[d] - self.halo[d]) for d in range(k)]
            others_hi = [range(0, self.shape[d]) for d in range(k + 1, dims)]
            for surface in (range(0, self.halo[k]),
                            range(self.shape[k] - self.halo[k], self.shape[k])):
                yield tuple(others_lo + [surface] + others_hi)

    def fixed_surface_iterator(self):
        """Points of the shell, slab by slab (ref ``halo_enumerator.py:46-76``).

        Within a slab the fixed dimension's index is the outermost loop and the remaining
        dimensions vary in ascending order, last fastest -- the reference's order.
        """
        for k, _ in enumerate(self.shape):
            for side in (0, 1):
                slab = list(self.slabs())[2 * k + side]
                rest = [slab[d] for d in range(len(slab)) if d != k]
                for fixed in slab[k]:
                    for other in itertools.product(*rest):
                        yield tuple(other[:k]) + (fixed,) + tuple(other[k:])

    def __iter__(self):
        return self.fixed_surface_iterator()

    def __len__(self):
        total = 0
        for slab in self.slabs():
            count = 1
            for r in slab:
                count *= len(r)
            total += count
        return total
