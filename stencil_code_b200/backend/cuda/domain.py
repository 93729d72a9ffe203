"""Iteration domain of one specialised launch.

For a single GPU the domain is the whole grid: indices are clamped to [0, n-1], the halo is the
``ghost_depth`` shell.  For one slab of a multi-GPU decomposition (axis 0 split, section 8e of
SURVEY.md) the local array carries ghost planes; a slab face that borders another slab is *open*:
no boundary handling happens there, the ghost planes hold the neighbour's data.  All four bounds
are compile-time constants of the generated kernel.
"""


class Domain(object):
    def __init__(self, shape, ghost_depth, open_lo=False, open_hi=False, slab_ghost=None, row_pitch=None):
        """
        :param shape: shape of the LOCAL array (ghost planes included when a face is open)
        :param ghost_depth: per-dimension ghost depth of the stencil
        :param open_lo / open_hi: axis-0 faces that border another slab
        :param slab_ghost: ghost planes per side of a slab (default: the stencil's axis-0 ghost
            depth; twice that when pairs of sweeps are fused between exchanges)
        :param row_pitch: elements between the starts of consecutive rows of every grid buffer
            (default: the row length).  iterate() keeps grids whose rows are no multiple of
            16 bytes in buffers with padded rows, which is what lets TMA address them.
        """
        self.shape = tuple(int(s) for s in shape)
        self.row_pitch = int(row_pitch) if row_pitch else self.shape[-1]
        if self.row_pitch < self.shape[-1]:
            raise ValueError("row pitch shorter than a row")
        self.ghost = tuple(int(g) for g in ghost_depth)
        self.open_lo, self.open_hi = bool(open_lo), bool(open_hi)
        dim = len(self.shape)
        h0 = self.ghost[0] if slab_ghost is None else int(slab_ghost)
        self.slab_ghost = h0
        # first / last index that holds real data of this rank's part of the global grid
        self.data_lo = [0] * dim
        self.data_hi = [n - 1 for n in self.shape]
        self.data_lo[0] = h0                      # slabs always carry ghost planes on both sides
        self.data_hi[0] = self.shape[0] - 1 - h0
        if not (open_lo or open_hi):
            self.data_lo[0], self.data_hi[0] = 0, self.shape[0] - 1
        # clamp bounds: global faces clamp to the first/last real plane, open faces never bind
        self.clamp_lo = list(self.data_lo)
        self.clamp_hi = list(self.data_hi)
        if open_lo:
            self.clamp_lo[0] = 0
        if open_hi:
            self.clamp_hi[0] = self.shape[0] - 1
        # interior (non-halo) index range [interior_lo, interior_hi)
        self.interior_lo = [self.data_lo[d] + self.ghost[d] for d in range(dim)]
        self.interior_hi = [self.data_hi[d] + 1 - self.ghost[d] for d in range(dim)]
        if open_lo:
            self.interior_lo[0] = 0
        if open_hi:
            self.interior_hi[0] = self.shape[0]

    @property
    def is_slab(self):
        return self.open_lo or self.open_hi

    @property
    def sweep_range(self):
        """Axis-0 index range [begin, end) a full sweep writes."""
        return self.data_lo[0], self.data_hi[0] + 1

    def key(self):
        return (self.shape, self.ghost, self.open_lo, self.open_hi, self.slab_ghost, self.row_pitch)

    @property
    def is_pitched(self):
        return self.row_pitch != self.shape[-1]
