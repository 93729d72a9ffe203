"""Slab-decomposed iterated stencils across the GPUs of one box.

The reference has no multi-device path at all; this is the north-star's addition for iterated
sweeps (Jacobi-3D).  One process per GPU.  The global grid is cut along axis 0 into contiguous
slabs; every rank's local array carries ``h`` ghost planes on both sides (``h`` = ghost depth along
axis 0).  A sweep on one rank is three launches of the SAME generated kernel over axis-0
sub-ranges (1-3 on an edge stream, 4 on the main stream, overlapping; see ``iterate``):

  1. top boundary planes    -> local result AND, through a peer-mapped pointer, straight into the
  2. bottom boundary planes    neighbour's ghost planes (NVLink stores issued by the stencil kernel
                               itself: no pack/copy/unpack, no NCCL in the data path)
  3. ``sc_signal``: system-scope fence + flag store (sweep number) into each neighbour's memory
  4. interior planes (the bulk; overlaps the neighbours' progress)

and the next sweep starts with a stream-ordered wait (``sc_stream_wait_value32``) until both
neighbours' flags reached the previous sweep number.  Interior planes never read ghost planes, so
a neighbour can never overwrite ghosts that are still being read (see ``iterate``).

``torch.distributed`` is used for the plumbing only: exchanging CUDA IPC handles, barriers and the
final reductions.  ``STENCIL_CUDA_EXCHANGE=nccl`` selects the conventional alternative -- the
boundary launches store locally only and the ghost planes travel by grouped NCCL send/recv on the
edge stream, still overlapped with the interior launch -- kept as the baseline the peer stores are
measured against (and for boxes without peer access between the GPUs).  Results are bit-identical to the single-GPU sweep: the per-point arithmetic is
the same generated code, only ``Domain`` (open faces) differs.
"""
import ctypes
import os

import numpy

from ...stencil_exception import StencilException


class SlabDecomposition(object):
    """Pure host arithmetic of the axis-0 slab cut (unit-tested on CPU with gloo)."""

    def __init__(self, global_shape, ghost0, world, rank, ghost_planes=None):
        self.global_shape = tuple(int(s) for s in global_shape)
        self.h = int(ghost0)                                    # the stencil's axis-0 ghost depth
        self.g = self.h if ghost_planes is None else int(ghost_planes)   # ghost planes per slab side
        self.world, self.rank = int(world), int(rank)
        n0 = self.global_shape[0]
        if self.world < 1 or not (0 <= self.rank < self.world):
            raise StencilException("bad rank {} of {}".format(rank, world))
        base, extra = divmod(n0, self.world)
        counts = [base + (1 if r < extra else 0) for r in range(self.world)]
        if min(counts) < max(1, 2 * self.g):
            raise StencilException("{} planes over {} ranks leaves slabs thinner than twice the "
                                   "{} ghost planes".format(n0, world, self.g))
        self.counts = counts
        self.starts = [sum(counts[:r]) for r in range(self.world)]
        self.start = self.starts[self.rank]
        self.planes = counts[self.rank]
        self.stop = self.start + self.planes
        self.open_lo = self.world > 1 and self.rank > 0
        self.open_hi = self.world > 1 and self.rank < self.world - 1
        self.ghosted = self.world > 1
        pad = 2 * self.g if self.ghosted else 0
        self.local_shape = (self.planes + pad,) + self.global_shape[1:]
        self.plane_elems = int(numpy.prod(self.global_shape[1:], dtype=numpy.int64)) if len(self.global_shape) > 1 else 1

    @property
    def first_real(self):
        """local axis-0 index of this rank's first real plane"""
        return self.g if self.ghosted else 0

    def local_index(self, global_plane):
        return global_plane - self.start + self.first_real

    def ranges(self):
        """(top, interior, bottom) local axis-0 ranges of one sweep; top/bottom are the planes a
        neighbour needs (and the only ones that read ghost planes)."""
        lo, hi = self.first_real, self.first_real + self.planes
        top = (lo, min(lo + self.g, hi)) if self.open_lo else (lo, lo)
        bottom = (max(hi - self.g, top[1]), hi) if self.open_hi else (hi, hi)
        return top, (top[1], bottom[0]), bottom

    def host_window(self):
        """Global axis-0 plane range [lo, hi) this rank needs from the host for one sweep: its real
        planes plus the ghost planes that border another slab."""
        lo = max(0, self.start - self.first_real)
        hi = min(self.global_shape[0], self.stop + (self.g if self.ghosted else 0))
        return lo, hi

    def sends(self):
        """[(neighbour rank, local source range, neighbour's local destination range)]"""
        plan = []
        lo, hi = self.first_real, self.first_real + self.planes
        if self.open_lo:
            up_planes = self.counts[self.rank - 1]
            plan.append((self.rank - 1, (lo, lo + self.g), (self.g + up_planes, 2 * self.g + up_planes)))
        if self.open_hi:
            plan.append((self.rank + 1, (hi - self.g, hi), (0, self.g)))
        return plan


class SlabStencil(object):
    """An iterated stencil on this rank's slab, exchanging ghost planes through peer memory."""

    def __init__(self, stencil, global_shape, tables=(), world=None, rank=None, device_index=None,
                 allow_fused=True):
        import torch.distributed as dist
        from . import runtime
        self.dist = dist
        self.runtime = runtime
        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.stencil = stencil
        self.device = runtime.Device.get(device_index)
        self.library = runtime.load_library()
        self.tables = [numpy.ascontiguousarray(t) for t in tables]
        h = stencil.ghost_depth[0]

        class _Shape(object):
            def __init__(self, shape, dtype):
                self.shape, self.dtype = shape, numpy.dtype(dtype)

        def specialise(slab):
            fake_args = (_Shape(slab.local_shape, numpy.float32),) + tuple(self.tables)
            return stencil.specializer.specialize(fake_args, open_faces=(slab.open_lo, slab.open_hi),
                                                  slab_ghost=slab.g if slab.ghosted else None)
        # Preferred layout: two ghost depths per side, so that pairs of sweeps run as one fused
        # launch (codegen_temporal) between ghost exchanges.  Falls back to one ghost depth and
        # single sweeps when the stencil is not eligible or the slabs are too thin.
        self.slab, self.concrete, self.fused = None, None, None
        if world > 1 and h > 0 and allow_fused:
            try:
                candidate = SlabDecomposition(global_shape, h, world, rank, ghost_planes=2 * h)
                concrete = specialise(candidate)
                if concrete.temporal() is not None:
                    self.slab, self.concrete, self.fused = candidate, concrete, concrete.temporal()
            except StencilException:
                pass
        if world > 1:
            # every rank must take the same decision (it fixes the ghost layout)
            votes = [None] * world
            dist.all_gather_object(votes, self.fused is not None)
            if not all(votes):
                self.slab, self.concrete, self.fused = None, None, None
        if self.slab is None:
            self.slab = SlabDecomposition(global_shape, h, world, rank)
            self.concrete = specialise(self.slab)
            if world == 1 and allow_fused:
                self.fused = self.concrete.temporal()
        slab = self.slab
        self.nbytes = int(numpy.prod(slab.local_shape, dtype=numpy.int64)) * 4
        self.buffers = [self.device.alloc(self.nbytes), self.device.alloc(self.nbytes)]
        self.flags = self.device.alloc(256)
        runtime.check(self.library.sc_memset32_async(ctypes.c_void_p(self.flags.ptr), 0, 64, self.device.stream))
        self.table_buffers = []
        for table in self.tables:
            buffer = self.device.alloc(table.nbytes)
            runtime.check(self.library.sc_memcpy_h2d_async(
                ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(table.ctypes.data), table.nbytes, self.device.stream))
            self.table_buffers.append(buffer)
        self.device.synchronize()
        self.sweep = 0          # sweeps applied so far
        self.step = 0           # launches-with-exchange so far (a fused step applies two sweeps)
        self.edge_stream = self.device._new_stream()
        self.edge_done = self.device.new_event()
        self.interior_done = self.device.new_event()
        self._events_armed = False
        self._ghosts_stale = False
        self._step_blocks = {}
        self.peers = {}
        self.exchange = os.environ.get("STENCIL_CUDA_EXCHANGE", "peer") if slab.ghosted else "none"
        if self.exchange not in ("peer", "nccl", "none"):
            raise StencilException("STENCIL_CUDA_EXCHANGE must be 'peer' or 'nccl'")
        if self.exchange == "peer":
            self._open_peers()
        elif self.exchange == "nccl":
            self._prepare_nccl()

    # -- setup -----------------------------------------------------------------------------------
    def _handle(self, pointer):
        raw = ctypes.create_string_buffer(64)
        self.runtime.check(self.library.sc_ipc_get_mem_handle(ctypes.c_void_p(pointer), raw))
        return raw.raw

    def _open(self, handle):
        pointer = ctypes.c_void_p()
        self.runtime.check(self.library.sc_ipc_open_mem_handle(handle, ctypes.byref(pointer)), "peer mapping")
        return pointer.value

    def _open_peers(self):
        mine = {"rank": self.slab.rank, "buffers": [self._handle(b.ptr) for b in self.buffers],
                "flags": self._handle(self.flags.ptr)}
        everyone = [None] * self.slab.world
        self.dist.all_gather_object(everyone, mine)
        for neighbour in (self.slab.rank - 1, self.slab.rank + 1):
            if 0 <= neighbour < self.slab.world:
                info = everyone[neighbour]
                self.peers[neighbour] = {"buffers": [self._open(h) for h in info["buffers"]],
                                         "flags": self._open(info["flags"])}
        self.dist.barrier()

    def _prepare_nccl(self):
        """torch views of the planes each buffer sends and receives, and the edge stream as a torch
        stream: the send/recv group is enqueued there, between the boundary launches and the
        event the next sweep's interior launch waits for."""
        import torch
        slab = self.slab
        plane = slab.plane_elems

        class _Planes(object):
            """zero-copy view of device memory for torch.as_tensor"""
            def __init__(self, pointer, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (pointer, False),
                                                 "version": 2}
        device = torch.device("cuda", self.device.index)

        def view(buffer, first, last):
            return torch.as_tensor(_Planes(buffer.ptr + first * plane * 4, (last - first) * plane), device=device)
        lo, hi = slab.first_real, slab.first_real + slab.planes
        self._nccl_views = []
        for buffer in self.buffers:
            views = {}
            if slab.open_lo:
                views["up"] = (view(buffer, lo, lo + slab.g), view(buffer, 0, slab.g))                  # send, recv
            if slab.open_hi:
                views["down"] = (view(buffer, hi - slab.g, hi), view(buffer, hi, hi + slab.g))
            self._nccl_views.append(views)
        self._edge_torch_stream = torch.cuda.ExternalStream(self.edge_stream.value if hasattr(self.edge_stream, "value")
                                                            else int(self.edge_stream), device=device)
        for neighbour in (slab.rank - 1, slab.rank + 1):
            if 0 <= neighbour < slab.world:
                self.peers[neighbour] = {"nccl": True}

    def _exchange_nccl(self, target_index):
        import torch
        dist, slab = self.dist, self.slab
        views = self._nccl_views[target_index]
        operations = []
        if "up" in views:
            send, receive = views["up"]
            operations += [dist.P2POp(dist.isend, send, slab.rank - 1), dist.P2POp(dist.irecv, receive, slab.rank - 1)]
        if "down" in views:
            send, receive = views["down"]
            operations += [dist.P2POp(dist.isend, send, slab.rank + 1), dist.P2POp(dist.irecv, receive, slab.rank + 1)]
        with torch.cuda.stream(self._edge_torch_stream):
            for work in dist.batch_isend_irecv(operations):
                work.wait()                      # stream-ordered: the edge stream waits, not the host

    def quiesce(self):
        """All ranks have finished every enqueued sweep: no neighbour kernel can still be storing
        into this rank's ghost planes (they are written through peer mappings, outside this
        rank's streams)."""
        self.synchronize()
        if self.slab.ghosted:
            self.dist.barrier()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if getattr(self, "_closed", False):
            return
        self._closed = True
        try:
            self.quiesce()
        except Exception:
            pass
        if self.exchange != "peer":
            self.peers = {}
            return
        for peer in self.peers.values():
            for pointer in peer["buffers"] + [peer["flags"]]:
                self.library.sc_ipc_close_mem_handle(ctypes.c_void_p(pointer))
        self.peers = {}

    # -- data in / out ----------------------------------------------------------------------------
    def _current(self):
        return self.buffers[self.step % 2]

    def fill_uniform(self, seed=1, scale=1.0):
        """Counter-based uniform [0, scale) fill, a function of the GLOBAL element index, so every
        decomposition of the same grid holds the same values (ghost planes included)."""
        slab = self.slab
        self.quiesce()
        self._ghosts_stale = False
        first_global = slab.start - slab.first_real            # global plane of local plane 0
        count = int(numpy.prod(slab.local_shape, dtype=numpy.int64))
        offset = first_global * slab.plane_elems               # may be negative on rank 0: wraps, unused
        geometry = type("L", (), dict(grid=(148 * 8, 1, 1), block=(256, 1, 1), cluster=(1, 1, 1), smem=0))()
        self.runtime.launch(self.device.utils(), "sc_fill_uniform", geometry,
                            [ctypes.c_void_p(self._current().ptr), ctypes.c_uint64(count), ctypes.c_uint64(seed),
                             ctypes.c_uint64(offset % (1 << 64)), ctypes.c_float(scale)], self.device.stream)
        self._resync_flags()

    def host_window(self):
        """Global axis-0 plane range [lo, hi) this rank needs from the host: its real planes plus
        the ghost planes that border another slab."""
        return self.slab.host_window()

    def apply_host(self, window, out=None):
        """ONE sweep, host memory to host memory, on this rank's slab: ``window`` holds the global
        planes ``host_window()`` (C-contiguous float32, ideally pinned); returns this rank's real
        planes of the result as a (pinned) host array.  The ghost planes come from the host
        window, so no exchange between the GPUs is needed; copy-in, sweep and copy-out are
        pipelined in axis-0 chunks on three streams (``ConcreteCudaStencil.sweep_host``).
        The device buffers keep the sweep's result: ``iterate()`` may follow."""
        slab = self.slab
        self.quiesce()
        lo, hi = self.host_window()
        window = numpy.ascontiguousarray(window, dtype=numpy.float32)
        if window.shape != (hi - lo,) + slab.global_shape[1:]:
            raise StencilException("apply_host needs planes {}..{} of the global grid".format(lo, hi))
        shape = (slab.planes,) + slab.global_shape[1:]
        if out is None:
            out = self.device.pinned_empty(shape, numpy.float32)
        elif out.shape != shape or out.dtype != numpy.float32 or not out.flags.c_contiguous:
            raise StencilException("out= must be a C-contiguous float32 array of this rank's real planes")
        local_lo = lo - (slab.start - slab.first_real)
        current = self.buffers[self.step % 2].ptr
        target = self.buffers[(self.step + 1) % 2].ptr
        first = slab.first_real
        plane_bytes = slab.plane_elems * 4
        self.concrete.sweep_host(window.ctypes.data, (local_lo, local_lo + (hi - lo)),
                                 [current] + [b.ptr for b in self.table_buffers], target,
                                 out.ctypes.data - first * plane_bytes, (first, first + slab.planes))
        self.step += 1
        self.sweep += 1
        if slab.ghosted:
            # the target buffer's ghost planes were not refreshed: iterate() must not trust them
            self._ghosts_stale = True
        return out

    def _refresh_ghosts(self):
        """Bring the ghost planes of the current buffer up to date from the neighbours' real
        planes (after ``apply_host``, which computes real planes only)."""
        slab = self.slab
        index = self.step % 2
        self.quiesce()
        if self.exchange == "nccl":
            self._exchange_nccl(index)
        else:
            plane_bytes = slab.plane_elems * 4
            for neighbour, source, destination in slab.sends():
                self.runtime.check(self.library.sc_memcpy_d2d_async(
                    ctypes.c_void_p(self.peers[neighbour]["buffers"][index] + destination[0] * plane_bytes),
                    ctypes.c_void_p(self.buffers[index].ptr + source[0] * plane_bytes),
                    (source[1] - source[0]) * plane_bytes, self.edge_stream))
        self._resync_flags()
        self._ghosts_stale = False

    def _resync_flags(self):
        """Every rank is about to be quiescent with up-to-date ghost planes: set this rank's flag
        words to the number of steps taken so far (sweeps that were no exchange steps -- apply_host
        -- advanced the step count without anybody publishing it)."""
        self.quiesce()
        if self.slab.ghosted and self.exchange == "peer":
            for offset in (0, 4):
                self.runtime.check(self.library.sc_stream_write_value32(
                    self.device.stream, ctypes.c_void_p(self.flags.ptr + offset), self.step))
            self.runtime.check(self.library.sc_memset32_async(
                ctypes.c_void_p(self.flags.ptr + 8), 0, 1, self.device.stream))
            self.quiesce()

    def load(self, global_grid):
        """Copy this rank's planes (plus ghost planes) of a host array of the global shape."""
        self.quiesce()
        slab = self.slab
        lo = max(0, slab.start - slab.first_real)
        hi = min(slab.global_shape[0], slab.stop + (slab.g if slab.ghosted else 0))
        part = numpy.ascontiguousarray(global_grid[lo:hi], dtype=numpy.float32)
        local_lo = lo - (slab.start - slab.first_real)
        destination = self._current().ptr + local_lo * slab.plane_elems * 4
        self.runtime.check(self.library.sc_memcpy_h2d_async(
            ctypes.c_void_p(destination), ctypes.c_void_p(part.ctypes.data), part.nbytes, self.device.stream))
        self._resync_flags()
        self._ghosts_stale = False

    def launches_per_step(self):
        """kernel launches one ``iterate`` step enqueues on this rank (a step = one sweep, or a
        fused pair of sweeps)"""
        if not self.slab.ghosted or self._step_kernel_available():
            return 1
        edges = int(self.slab.open_lo) + int(self.slab.open_hi)
        return 1 + edges + (1 if self.exchange == "peer" else 0)

    def result(self, out=None):
        """This rank's real planes as a host array."""
        self.synchronize()
        slab = self.slab
        shape = (slab.planes,) + slab.global_shape[1:]
        if out is None:
            out = self.device.pinned_empty(shape, numpy.float32)
        elif out.shape != shape or out.dtype != numpy.float32 or not out.flags.c_contiguous:
            raise StencilException("out= must be a C-contiguous float32 array of this rank's real planes")
        source = self._current().ptr + slab.first_real * slab.plane_elems * 4
        self.runtime.check(self.library.sc_memcpy_d2h_async(
            ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(source), out.nbytes, self.device.stream))
        self.device.synchronize()
        return out

    def checksum(self):
        """Position-dependent 64-bit checksum of the real planes; summed (mod 2^64) over ranks it
        is independent of the decomposition."""
        self.synchronize()
        slab = self.slab
        count = slab.planes * slab.plane_elems
        cell = self.device.alloc(256)
        self.runtime.check(self.library.sc_memset32_async(ctypes.c_void_p(cell.ptr), 0, 2, self.device.stream))
        geometry = type("L", (), dict(grid=(148 * 8, 1, 1), block=(256, 1, 1), cluster=(1, 1, 1), smem=0))()
        source = self._current().ptr + slab.first_real * slab.plane_elems * 4
        self.runtime.launch(self.device.utils(), "sc_checksum", geometry,
                            [ctypes.c_void_p(source), ctypes.c_uint64(count),
                             ctypes.c_uint64(slab.start * slab.plane_elems), ctypes.c_void_p(cell.ptr)],
                            self.device.stream)
        host = numpy.zeros(1, numpy.uint64)
        self.runtime.check(self.library.sc_memcpy_d2h_async(
            ctypes.c_void_p(host.ctypes.data), ctypes.c_void_p(cell.ptr), 8, self.device.stream))
        self.device.synchronize()
        cell.release()
        value = int(host[0])
        if slab.ghosted:
            parts = [None] * slab.world
            self.dist.all_gather_object(parts, value)
            value = sum(parts) % (1 << 64)
        return value

    # -- sweeps ------------------------------------------------------------------------------------
    def iterate(self, sweeps):
        """Enqueue ``sweeps`` sweeps (asynchronous; see ``synchronize``).

        Two streams per rank: the *edge* stream runs the boundary-plane launches (which also store
        into the neighbours' ghost planes) and the flag signal, the main stream runs the interior
        launch; both read the current buffer and write disjoint planes of the next one, so they
        overlap.  Hazards between sweeps are closed by two events and the neighbours' flags:

          * edge(t) waits interior(t-1): it reads planes the interior wrote, and writes planes of
            the buffer interior(t-1) was still reading;
          * interior(t) waits edge(t-1): same two reasons, mirrored;
          * edge(t) waits for both neighbours' flags >= t-1: their boundary results of sweep t-1
            are in my ghost planes, and they are done reading the ghost planes I am about to
            overwrite (ghost planes are only ever read by edge kernels).
        """
        slab = self.slab
        runtime, library = self.runtime, self.library
        main, edge = self.device.stream, self.edge_stream
        tables = [b.ptr for b in self.table_buffers]
        if self._ghosts_stale and sweeps > 0:
            self._refresh_ghosts()
        if not slab.ghosted:
            # one GPU: the same enqueue-from-C ping-pong (programmatic dependent launch, fused
            # pairs of sweeps) that ``stencil.iterate`` uses
            current, spare = self.buffers[self.step % 2].ptr, self.buffers[(self.step + 1) % 2].ptr
            final = self.concrete.sweeps(current, spare, tables, sweeps, main, allow_fused=self.fused is not None)
            if final != current:
                self.step += 1
            self.sweep += sweeps
            return
        top, interior, bottom = slab.ranges()
        up, down = self.peers.get(slab.rank - 1), self.peers.get(slab.rank + 1)
        plane = slab.plane_elems
        signal_geometry = type("L", (), dict(grid=(1, 1, 1), block=(32, 1, 1), cluster=(1, 1, 1), smem=0))()
        exchanging = bool(up or down)
        remaining = sweeps
        if exchanging and remaining >= 2 and self._step_kernel_available():
            remaining -= 2 * self._enqueue_fused_steps(remaining // 2)
        while remaining > 0:
            fused = self.fused is not None and remaining >= 2
            applied = 2 if fused else 1

            def run(begin, end, stream, peer_pointer=0, peer_shift=0):
                if fused:
                    self.concrete._launch_plan(self.fused[0], self.fused[1], [current] + tables, target,
                                               begin, end, stream, peer_pointer, peer_shift)
                else:
                    self.concrete.launch([current] + tables, target, begin, end, stream,
                                         peer_pointer=peer_pointer, peer_shift=peer_shift)
            number = self.step + 1
            current = self.buffers[self.step % 2].ptr
            target_index = (self.step + 1) % 2
            target = self.buffers[target_index].ptr
            if exchanging:
                if number > 1 or self._events_armed:
                    runtime.check(library.sc_stream_wait_event(edge, self.interior_done))
                    runtime.check(library.sc_stream_wait_event(main, self.edge_done))
                if number > 1 and self.exchange == "peer":
                    # flags: [0] written by the upper neighbour, [1] by the lower one
                    if up:
                        runtime.check(library.sc_stream_wait_value32(edge, ctypes.c_void_p(self.flags.ptr), number - 1))
                    if down:
                        runtime.check(library.sc_stream_wait_value32(edge, ctypes.c_void_p(self.flags.ptr + 4), number - 1))
                if self.exchange == "nccl":
                    # boundary planes land locally; a grouped send/recv carries them into the
                    # neighbours' ghost planes of the same buffer (ordering between ranks is NCCL's)
                    if up:
                        run(top[0], top[1], edge)
                    if down:
                        run(bottom[0], bottom[1], edge)
                    self._exchange_nccl(target_index)
                else:
                    if up:
                        # my first g real planes become the upper neighbour's lower ghost planes
                        run(top[0], top[1], edge, up["buffers"][target_index], slab.counts[slab.rank - 1] * plane)
                    if down:
                        run(bottom[0], bottom[1], edge, down["buffers"][target_index], -slab.planes * plane)
                    runtime.launch(self.device.utils(), "sc_signal", signal_geometry,
                                   [ctypes.c_void_p(up["flags"] + 4 if up else 0),
                                    ctypes.c_void_p(down["flags"] if down else 0),
                                    ctypes.c_uint32(number)], edge)
                runtime.check(library.sc_event_record(self.edge_done, edge))
            run(interior[0], interior[1], main)
            if exchanging:
                runtime.check(library.sc_event_record(self.interior_done, main))
                self._events_armed = True
            self.step += 1
            self.sweep += applied
            remaining -= applied

    # -- one launch per exchange step ---------------------------------------------------------------
    def _step_kernel_available(self):
        return (self.fused is not None and self.exchange == "peer"
                and getattr(self.fused[0], "function_step", None) is not None
                and os.environ.get("STENCIL_CUDA_STEP_KERNEL", "1") not in ("0", ""))

    def _enqueue_fused_steps(self, count):
        """``count`` fused double sweeps, each ONE launch of ``stencil_temporal2_step``: its first CTAs
        compute the boundary planes next to each neighbour (waiting for the neighbour's flag, storing
        into the neighbour's ghost planes, publishing this rank's flag when all are done), the rest
        the interior.  All launches are enqueued by one C call (``sc_launch_steps``) on the main
        stream, with programmatic dependent launch; the GPUs synchronise among themselves through
        the flag words, the host is not involved."""
        runtime, library, slab = self.runtime, self.library, self.slab
        plan, module = self.fused
        programmatic = 0 if os.environ.get("STENCIL_CUDA_PDL", "1") in ("0", "") else 1
        if programmatic:
            module = self.concrete._dependent_launch_module(plan)
        main = self.device.stream
        if self._events_armed:                    # sweeps enqueued the two-stream way come first
            runtime.check(library.sc_stream_wait_event(main, self.edge_done))
        lo, hi = slab.first_real, slab.first_real + slab.planes
        geometry = plan.step_geometry(lo, hi)
        top, interior, bottom = geometry.ranges
        up, down = self.peers.get(slab.rank - 1), self.peers.get(slab.rank + 1)
        plane = slab.plane_elems
        blocks = []
        for parity in (self.step % 2, (self.step + 1) % 2):
            current, target_index = self.buffers[parity].ptr, (parity + 1) % 2
            key = ("step", parity)
            block = self._step_blocks.get(key)
            if block is None:
                values = [ctypes.c_void_p(current)] + [ctypes.c_void_p(b.ptr) for b in self.table_buffers]
                for spec in plan.tensor_maps:
                    values.append(self.concrete._tensor_map(spec, current))
                values += [ctypes.c_void_p(self.buffers[target_index].ptr),
                           ctypes.c_int(interior[0]), ctypes.c_int(interior[1]), ctypes.c_int(geometry.extra[0]),
                           ctypes.c_void_p(up["buffers"][target_index] if up else 0),
                           ctypes.c_longlong(slab.counts[slab.rank - 1] * plane if up else 0),
                           ctypes.c_void_p(down["buffers"][target_index] if down else 0),
                           ctypes.c_longlong(-slab.planes * plane if down else 0),
                           ctypes.c_int(top[0]), ctypes.c_int(top[1]), ctypes.c_int(bottom[0]), ctypes.c_int(bottom[1]),
                           ctypes.c_void_p(self.flags.ptr),
                           ctypes.c_void_p(up["flags"] + 4 if up else 0),        # I am the upper neighbour's lower one
                           ctypes.c_void_p(down["flags"] if down else 0),
                           ctypes.c_uint32(0)]
                pointers = (ctypes.c_void_p * len(values))(
                    *[ctypes.cast(ctypes.byref(v), ctypes.c_void_p) for v in values])
                block = (values, pointers)
                self._step_blocks[key] = block
            blocks.append(block)
        step_arg = len(blocks[0][0]) - 1
        grid = (ctypes.c_uint32 * 3)(*geometry.grid)
        threads = (ctypes.c_uint32 * 3)(*geometry.block)
        runtime.check(library.sc_launch_steps(module.handle, plan.function_step.encode(), grid, threads,
                                              geometry.smem, main, blocks[0][1], blocks[1][1], count,
                                              programmatic, step_arg, self.step), "fused exchange steps")
        runtime.check(library.sc_event_record(self.interior_done, main))
        runtime.check(library.sc_event_record(self.edge_done, main))
        self._events_armed = True
        self.step += count
        self.sweep += 2 * count
        return count

    def join(self):
        """Make the main stream wait for the edge stream's last sweep (for timing with events
        recorded on the main stream)."""
        if self._events_armed:
            self.runtime.check(self.library.sc_stream_wait_event(self.device.stream, self.edge_done))

    def synchronize(self):
        self.device.synchronize(self.device.stream)
        self.device.synchronize(self.edge_stream)
