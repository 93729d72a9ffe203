"""Host runtime of the CUDA backend: ctypes binding of the C-ABI shim and the concrete callable.

Replaces reference ``ConcreteStencil`` / ``OclStencilFunction`` (``stencil_kernel.py:61-208``):
``ConcreteCudaStencil.__call__`` is what ``Stencil.__call__`` ends up in.  numpy arrays go
host -> device -> kernel -> host (what ``OclStencilFunction.__call__`` does with pycl buffers,
``:162-202``); torch CUDA tensors are used in place and a CUDA tensor is returned (the analogue of
the reference's device-resident ``hmarray`` path, ``:153-165,197-198``).

There is no fallback: if ``libstencil_cuda.so`` or a GPU is missing the call raises
``StencilException``.
"""
import atexit
import ctypes
import os
import threading
import weakref

import numpy

from ...stencil_exception import StencilException

_CSRC = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "csrc"))
LIBRARY_PATH = os.path.join(_CSRC, "libstencil_cuda.so")

_lib = None
_lib_lock = threading.Lock()


class ScTmap(ctypes.Structure):
    _fields_ = [("opaque", ctypes.c_uint64 * 16)]
    _align_ = 64


class ScDeviceProps(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 128),
                ("cc_major", ctypes.c_int32), ("cc_minor", ctypes.c_int32),
                ("sm_count", ctypes.c_int32), ("max_threads_per_block", ctypes.c_int32),
                ("max_threads_per_sm", ctypes.c_int32), ("regs_per_sm", ctypes.c_int32),
                ("smem_per_sm", ctypes.c_int32), ("smem_per_block_optin", ctypes.c_int32),
                ("l2_bytes", ctypes.c_int32), ("sm_clock_khz", ctypes.c_int32),
                ("mem_clock_khz", ctypes.c_int32), ("reserved", ctypes.c_int32 * 4),
                ("total_mem", ctypes.c_uint64)]


# name -> (restype, argtypes); every symbol include/stencil_cuda.h declares
_VP, _I, _SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
_PVP = ctypes.POINTER(ctypes.c_void_p)
_U32x3 = ctypes.POINTER(ctypes.c_uint32)
SYMBOLS = {
    "sc_abi_version": (_I, []),
    "sc_last_error": (ctypes.c_char_p, []),
    "sc_driver_available": (_I, []),
    "sc_device_count": (_I, [ctypes.POINTER(_I)]),
    "sc_init": (_I, [_I]),
    "sc_set_device": (_I, [_I]),
    "sc_get_device": (_I, [ctypes.POINTER(_I)]),
    "sc_device_props_get": (_I, [_I, ctypes.POINTER(ScDeviceProps)]),
    "sc_mem_info": (_I, [ctypes.POINTER(_SZ), ctypes.POINTER(_SZ)]),
    "sc_device_synchronize": (_I, []),
    "sc_compile": (_I, [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), _I, _PVP]),
    "sc_module_from_cubin": (_I, [_VP, _SZ, _PVP]),
    "sc_module_load": (_I, [_VP]),
    "sc_module_cubin": (_I, [_VP, _PVP, ctypes.POINTER(_SZ)]),
    "sc_module_log": (ctypes.c_char_p, [_VP]),
    "sc_module_free": (_I, [_VP]),
    "sc_function_info": (_I, [_VP, ctypes.c_char_p] + [ctypes.POINTER(_I)] * 4),
    "sc_function_occupancy": (_I, [_VP, ctypes.c_char_p, _I, _SZ, ctypes.POINTER(_I)]),
    "sc_tensormap_tiled": (_I, [ctypes.POINTER(ScTmap), _VP, _I, _I, ctypes.POINTER(ctypes.c_uint64),
                                ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32), _I]),
    "sc_launch": (_I, [_VP, ctypes.c_char_p, _U32x3, _U32x3, _U32x3, ctypes.c_uint32, _VP, _PVP]),
    "sc_launch_pingpong": (_I, [_VP, ctypes.c_char_p, _U32x3, _U32x3, ctypes.c_uint32, _VP, _PVP, _PVP, _I, ctypes.c_uint32]),
    "sc_prepare_launch": (_I, [_VP, ctypes.c_char_p, _U32x3, _U32x3, _U32x3, ctypes.c_uint32, _PVP, _PVP]),
    "sc_launch_prepared": (_I, [_VP, _VP]),
    "sc_prepared_free": (_I, [_VP]),
    "sc_launch_steps": (_I, [_VP, ctypes.c_char_p, _U32x3, _U32x3, ctypes.c_uint32, _VP, _PVP, _PVP, _I, ctypes.c_uint32,
                             _I, ctypes.c_uint32]),
    "sc_stream_create": (_I, [_PVP]),
    "sc_stream_destroy": (_I, [_VP]),
    "sc_stream_synchronize": (_I, [_VP]),
    "sc_stream_wait_event": (_I, [_VP, _VP]),
    "sc_event_create": (_I, [_PVP, _I]),
    "sc_event_create_ipc": (_I, [_PVP]),
    "sc_event_destroy": (_I, [_VP]),
    "sc_event_record": (_I, [_VP, _VP]),
    "sc_event_synchronize": (_I, [_VP]),
    "sc_event_elapsed_ms": (_I, [_VP, _VP, ctypes.POINTER(ctypes.c_float)]),
    "sc_malloc": (_I, [_PVP, _SZ]),
    "sc_free": (_I, [_VP]),
    "sc_memset32_async": (_I, [_VP, ctypes.c_uint32, _SZ, _VP]),
    "sc_memcpy_h2d_async": (_I, [_VP, _VP, _SZ, _VP]),
    "sc_memcpy_d2h_async": (_I, [_VP, _VP, _SZ, _VP]),
    "sc_memcpy_d2d_async": (_I, [_VP, _VP, _SZ, _VP]),
    "sc_memcpy2d_async": (_I, [_VP, _SZ, _VP, _SZ, _SZ, _SZ, _VP]),
    "sc_host_alloc": (_I, [_PVP, _SZ]),
    "sc_host_free": (_I, [_VP]),
    "sc_host_register": (_I, [_VP, _SZ]),
    "sc_host_unregister": (_I, [_VP]),
    "sc_stream_wait_value32": (_I, [_VP, _VP, ctypes.c_uint32]),
    "sc_stream_write_value32": (_I, [_VP, _VP, ctypes.c_uint32]),
    "sc_ipc_get_mem_handle": (_I, [_VP, ctypes.c_char_p]),
    "sc_ipc_open_mem_handle": (_I, [ctypes.c_char_p, _PVP]),
    "sc_ipc_close_mem_handle": (_I, [_VP]),
    "sc_ipc_get_event_handle": (_I, [_VP, ctypes.c_char_p]),
    "sc_ipc_open_event_handle": (_I, [ctypes.c_char_p, _PVP]),
}


def load_library():
    """The shim, loaded once.  Raises StencilException when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIBRARY_PATH):
                raise StencilException(
                    "the CUDA shim {} is missing; build it with `make -C {}` (or "
                    "`python -c 'import __graft_entry__ as g; g.build()'`)".format(LIBRARY_PATH, _CSRC))
            library = ctypes.CDLL(LIBRARY_PATH)
            for name, (restype, argtypes) in SYMBOLS.items():
                function = getattr(library, name)
                function.restype = restype
                function.argtypes = argtypes
            if library.sc_abi_version() != 1:
                raise StencilException("libstencil_cuda.so has ABI version {}, expected 1".format(
                    library.sc_abi_version()))
            _lib = library
    return _lib


def check(code, what=""):
    if code != 0:
        message = load_library().sc_last_error().decode(errors="replace")
        raise StencilException("CUDA backend error{}: {}".format(
            " in " + what if what else "", message or "code {}".format(code)))


# ---- modules -------------------------------------------------------------------------------------
class Module(object):
    def __init__(self, handle, name):
        self.handle = handle
        self.name = name
        self.from_disk_cache = False
        self._keepalive = []

    def cubin(self):
        data, size = ctypes.c_void_p(), ctypes.c_size_t()
        check(load_library().sc_module_cubin(self.handle, ctypes.byref(data), ctypes.byref(size)))
        return ctypes.string_at(data, size.value)

    def log(self):
        return load_library().sc_module_log(self.handle).decode(errors="replace")

    def load(self):
        check(load_library().sc_module_load(self.handle), "module load")

    def function_info(self, function):
        values = [ctypes.c_int() for _ in range(4)]
        check(load_library().sc_function_info(self.handle, function.encode(),
                                              *[ctypes.byref(v) for v in values]))
        return dict(zip(("regs", "static_smem", "local_bytes", "max_threads"),
                        (v.value for v in values)))

    def occupancy(self, function, block_threads, dynamic_smem):
        blocks = ctypes.c_int()
        check(load_library().sc_function_occupancy(self.handle, function.encode(), block_threads,
                                                   dynamic_smem, ctypes.byref(blocks)))
        return blocks.value


def _cache_dir():
    """Directory of the on-disk cubin cache (the counterpart of ctree's on-disk JIT cache, which
    the reference relies on for its "without compile" timings); STENCIL_CUDA_CACHE=off disables."""
    configured = os.environ.get("STENCIL_CUDA_CACHE")
    if configured in ("off", "0"):
        return None
    path = configured or os.path.join(os.path.expanduser("~"), ".cache", "stencil_code_b200")
    try:
        os.makedirs(path, exist_ok=True)
        return path
    except OSError:
        return None


def compile_source(source, name="stencil_kernel.cu", options=()):
    """CUDA C++ -> sm_100a module.  Looks in the on-disk cubin cache first (keyed by source and
    options), else compiles with NVRTC through the shim (works without a GPU) and stores the cubin."""
    import hashlib
    library = load_library()
    directory = _cache_dir()
    cached_path = None
    if directory is not None:
        digest = hashlib.sha256(("\n".join(options) + "\0" + source).encode()).hexdigest()[:32]
        cached_path = os.path.join(directory, "{}_{}.cubin".format(os.path.splitext(name)[0], digest))
        if os.path.exists(cached_path):
            try:
                with open(cached_path, "rb") as handle:
                    data = handle.read()
                if data[:4] == b"\x7fELF":
                    module = module_from_cubin(data, name)
                    module.from_disk_cache = True
                    return module
            except (OSError, StencilException):
                pass
    raw = (ctypes.c_char_p * len(options))(*[o.encode() for o in options])
    handle = ctypes.c_void_p()
    check(library.sc_compile(source.encode(), name.encode(), raw, len(options), ctypes.byref(handle)),
          "JIT compile of " + name)
    module = Module(handle, name)
    if cached_path is not None:
        try:
            temporary = cached_path + ".tmp{}".format(os.getpid())
            with open(temporary, "wb") as out:
                out.write(module.cubin())
            os.replace(temporary, cached_path)
        except OSError:
            pass
    return module


def module_from_cubin(data, name="cubin"):
    handle = ctypes.c_void_p()
    buffer = ctypes.create_string_buffer(data, len(data))
    check(load_library().sc_module_from_cubin(buffer, len(data), ctypes.byref(handle)))
    return Module(handle, name)


# ---- devices ---------------------------------------------------------------------------------------
class Device(object):
    """One initialised GPU: properties, a compute stream, two copy streams, buffer caches."""
    _instances = {}
    _lock = threading.Lock()

    def __init__(self, index):
        library = load_library()
        if not library.sc_driver_available():
            check(library.sc_init(index), "device initialisation")      # raises with the reason
        check(library.sc_init(index), "device initialisation")
        self.index = index
        props = ScDeviceProps()
        check(library.sc_device_props_get(index, ctypes.byref(props)))
        self.props = props
        self.name = props.name.decode()
        self.stream = self._new_stream()
        self.h2d_stream = self._new_stream()
        self.d2h_stream = self._new_stream()
        self._free_device = {}       # nbytes -> [ptr]
        self._free_pinned = {}       # nbytes -> [ptr]
        self._pool_lock = threading.RLock()   # finalizers of dropped buffers may run on any thread
        self._events = []
        self._utils = None

    @classmethod
    def get(cls, index=None):
        if index is None:
            index = int(os.environ.get("STENCIL_CUDA_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        with cls._lock:
            device = cls._instances.get(index)
            if device is None:
                device = cls(index)
                cls._instances[index] = device
        check(load_library().sc_set_device(index))
        return device

    def planner_props(self):
        from .planner import DeviceProps
        p = self.props
        return DeviceProps(sm_count=p.sm_count, smem_per_sm=p.smem_per_sm,
                           smem_per_block_optin=p.smem_per_block_optin, regs_per_sm=p.regs_per_sm,
                           max_threads_per_sm=p.max_threads_per_sm, l2_bytes=p.l2_bytes,
                           name=self.name)

    # -- streams / events
    def _new_stream(self):
        stream = ctypes.c_void_p()
        check(load_library().sc_stream_create(ctypes.byref(stream)))
        return stream

    def event(self):
        """A pooled non-timing event."""
        if self._events:
            return self._events.pop()
        return self.new_event(False)

    def recycle_event(self, event):
        self._events.append(event)

    def new_event(self, timing=False):
        event = ctypes.c_void_p()
        check(load_library().sc_event_create(ctypes.byref(event), 1 if timing else 0))
        return event

    def synchronize(self, stream=None):
        if stream is None:
            check(load_library().sc_device_synchronize())
        else:
            check(load_library().sc_stream_synchronize(stream))

    # -- device memory (size-keyed free lists: a call's buffers are reused by the next call)
    def alloc(self, nbytes):
        nbytes = max(256, (int(nbytes) + 255) // 256 * 256)
        with self._pool_lock:
            bucket = self._free_device.get(nbytes)
            pooled = bucket.pop() if bucket else None
        if pooled is not None:
            return DeviceBuffer(self, pooled, nbytes)
        pointer = ctypes.c_void_p()
        code = load_library().sc_malloc(ctypes.byref(pointer), nbytes)
        if code != 0:
            self.trim()
            check(load_library().sc_malloc(ctypes.byref(pointer), nbytes), "device allocation")
        return DeviceBuffer(self, pointer.value, nbytes)

    def _release(self, pointer, nbytes):
        with self._pool_lock:
            self._free_device.setdefault(nbytes, []).append(pointer)

    def trim(self):
        """Return every cached device / pinned block to the driver."""
        library = load_library()
        with self._pool_lock:
            for bucket in self._free_device.values():
                while bucket:
                    library.sc_free(ctypes.c_void_p(bucket.pop()))
            for bucket in self._free_pinned.values():
                while bucket:
                    library.sc_host_free(ctypes.c_void_p(bucket.pop()))

    # -- pinned host memory
    def pinned_empty(self, shape, dtype):
        """numpy array in page-locked memory; the block returns to a pool when the array dies."""
        dtype = numpy.dtype(dtype)
        count = int(numpy.prod(shape, dtype=numpy.int64)) if len(shape) else 1
        nbytes = max(4096, (count * dtype.itemsize + 4095) // 4096 * 4096)
        with self._pool_lock:
            bucket = self._free_pinned.get(nbytes)
            pointer = bucket.pop() if bucket else None
        if pointer is None:
            raw = ctypes.c_void_p()
            check(load_library().sc_host_alloc(ctypes.byref(raw), nbytes), "pinned allocation")
            pointer = raw.value
        block = (ctypes.c_char * nbytes).from_address(pointer)
        array = numpy.frombuffer(block, dtype=dtype, count=count).reshape(shape)
        weakref.finalize(block, self._recycle_pinned, pointer, nbytes)
        return array

    PINNED_POOL_LIMIT = 4 << 30         # bytes of idle page-locked blocks kept for reuse

    def _recycle_pinned(self, pointer, nbytes):
        """An array from pinned_empty died: keep its block for the next call of the same size.
        The idle pool is bounded: beyond PINNED_POOL_LIMIT older idle blocks go back to the
        driver, so that at most the limit -- or one single larger block -- stays page-locked."""
        library = load_library()
        with self._pool_lock:
            pool = self._free_pinned
            idle = sum(size * len(bucket) for size, bucket in pool.items())
            for size in list(pool):
                bucket = pool[size]
                while bucket and idle + nbytes > self.PINNED_POOL_LIMIT:
                    library.sc_host_free(ctypes.c_void_p(bucket.pop(0)))
                    idle -= size
                if not bucket:
                    del pool[size]
            pool.setdefault(nbytes, []).append(pointer)

    # -- helper kernels (csrc/device_utils.cu)
    def utils(self):
        if self._utils is None:
            cubin_path = os.path.join(_CSRC, "device_utils.cubin")
            if os.path.exists(cubin_path):
                with open(cubin_path, "rb") as handle:
                    self._utils = module_from_cubin(handle.read(), "device_utils")
            else:
                with open(os.path.join(_CSRC, "device_utils.cu")) as handle:
                    self._utils = compile_source(handle.read(), "device_utils.cu", ("-lineinfo",))
            self._utils.load()
        return self._utils


def pinned_empty(shape, dtype=numpy.float32, device_index=None):
    """Uninitialised numpy array in page-locked host memory (fast, asynchronous transfers)."""
    if isinstance(shape, int):
        shape = (shape,)
    return Device.get(device_index).pinned_empty(tuple(shape), dtype)


class DeviceBuffer(object):
    """A device allocation that goes back to its device's free list when dropped."""
    def __init__(self, device, pointer, nbytes):
        self.device, self.ptr, self.nbytes = device, pointer, nbytes
        self._finalizer = weakref.finalize(self, device._release, pointer, nbytes)

    def release(self):
        self._finalizer()


@atexit.register
def _shutdown():
    for device in list(Device._instances.values()):
        try:
            load_library().sc_set_device(device.index)
            device.trim()
        except Exception:
            pass


def launch(module, function, launch_cfg, args, stream):
    """args: list of ctypes objects (by value); marshalled as an array of pointers to them."""
    library = load_library()
    pointers = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.byref(a), ctypes.c_void_p) for a in args])
    grid = (ctypes.c_uint32 * 3)(*launch_cfg.grid)
    block = (ctypes.c_uint32 * 3)(*launch_cfg.block)
    cluster = (ctypes.c_uint32 * 3)(*launch_cfg.cluster)
    check(library.sc_launch(module.handle, function.encode(), grid, block, cluster,
                            launch_cfg.smem, stream, pointers), "launch of " + function)


class _PreparedLaunch(object):
    """A launch resolved once (``sc_prepare_launch``): kernel handle, shared-memory opt-in, geometry
    and the argument block; freed with the cache entry that holds it."""
    def __init__(self, module, name, grid, block, cluster, smem, pointers):
        self.handle = ctypes.c_void_p()
        self._keep = (module, pointers, grid, block, cluster)
        check(load_library().sc_prepare_launch(module.handle, name, grid, block, cluster, smem, pointers,
                                               ctypes.byref(self.handle)), "preparing a launch")

    def __del__(self):
        try:
            if self.handle:
                load_library().sc_prepared_free(self.handle)
        except Exception:
            pass


# ---- the concrete callable ----------------------------------------------------------------------
def _is_torch_tensor(obj):
    return hasattr(obj, "data_ptr") and hasattr(obj, "is_cuda")


def _device_tensors(args):
    """contiguous tensors on the grid's device for every argument (tables may arrive as numpy)"""
    import torch
    device = args[0].device
    return [(arg if _is_torch_tensor(arg) else torch.as_tensor(arg)).to(device).contiguous() for arg in args]


class ConcreteCudaStencil(object):
    """A compiled, launchable specialisation (what ``SpecializedStencil.finalize`` returns)."""

    def __init__(self, plan, device_index=None):
        self.plan = plan
        self.program = plan.program
        self.device = Device.get(device_index)
        self.module = compile_source(plan.source, "stencil_{}.cu".format(plan.strategy), plan.options)
        self.module.load()
        self._tmaps = {}
        self._launches = {}
        self._temporal = None
        self._pdl_modules = {}           # id(plan) -> module compiled with -DSC_PDL
        self._twin = None                # specialisation over padded rows (see _pitched_twin)
        self._pairs = {}                 # id(other concrete stencil) -> (other, fused (plan, module) or None)
        self._sc_launch = load_library().sc_launch
        self._sc_launch_prepared = load_library().sc_launch_prepared
        self.last_duration = None        # seconds of the last call when STENCIL_CUDA_TIMING=1

    # -- argument checks (the reference's ndpointer argtypes raise TypeError on mismatch)
    def _check_args(self, args):
        program = self.program
        if len(args) != len(program.args):
            raise StencilException("stencil takes {} arguments, got {}".format(len(program.args), len(args)))
        for info, arg in zip(program.args, args):
            shape = tuple(int(s) for s in arg.shape)
            dtype = numpy.dtype(str(arg.dtype).replace("torch.", "")) if _is_torch_tensor(arg) \
                else numpy.dtype(arg.dtype)
            if shape != info.shape or dtype != info.dtype:
                raise TypeError("argument '{}' must have shape {} and dtype {}, got {} {}".format(
                    info.name, info.shape, info.dtype, shape, dtype))

    def _tensor_map(self, spec, pointer):
        key = (spec.key, pointer)
        tmap = self._tmaps.get(key)
        if tmap is None:
            tmap = ScTmap()
            rank = len(spec.dims)
            dims = (ctypes.c_uint64 * rank)(*spec.dims)
            strides = (ctypes.c_uint64 * max(1, rank - 1))(*spec.strides_bytes)
            box = (ctypes.c_uint32 * rank)(*spec.box)
            check(load_library().sc_tensormap_tiled(ctypes.byref(tmap), ctypes.c_void_p(pointer),
                                                    spec.dtype_code, rank, dims, strides, box, 0),
                  "tensor map encode")
            while len(self._tmaps) >= 256:          # least recently used goes first
                self._tmaps.pop(next(iter(self._tmaps)))
        else:
            del self._tmaps[key]                    # re-insert: dicts keep insertion order
        self._tmaps[key] = tmap
        return tmap

    def launch(self, arg_pointers, out_pointer, a0_begin=None, a0_end=None, stream=None,
               peer_pointer=0, peer_shift=0):
        """Enqueue the stencil over axis-0 range [a0_begin, a0_end) on ``stream``.  With
        ``peer_pointer`` every result is also stored at ``peer_pointer[flat_index + peer_shift]``
        (a neighbouring slab's ghost planes, mapped peer memory).  The marshalled argument block
        of a (pointers, range) combination is cached: repeated sweeps cost one C call."""
        self._launch_plan(self.plan, self.module, arg_pointers, out_pointer, a0_begin, a0_end, stream,
                          peer_pointer, peer_shift)

    def _launch_plan(self, plan, module, arg_pointers, out_pointer, a0_begin, a0_end, stream,
                     peer_pointer=0, peer_shift=0, prepare_only=False):
        lo, hi = plan.domain.sweep_range
        a0_begin = lo if a0_begin is None else a0_begin
        a0_end = hi if a0_end is None else a0_end
        if a0_end <= a0_begin:
            return
        key = (id(plan), tuple(arg_pointers), out_pointer, a0_begin, a0_end, peer_pointer, peer_shift)
        prepared = self._launches.get(key)
        if prepared is None:
            values = [ctypes.c_void_p(pointer) for pointer in arg_pointers]
            for spec in plan.tensor_maps:
                values.append(self._tensor_map(spec, arg_pointers[spec.arg]))
            values.append(ctypes.c_void_p(out_pointer))
            values.append(ctypes.c_void_p(peer_pointer))
            values.append(ctypes.c_longlong(peer_shift))
            values.append(ctypes.c_int(a0_begin))
            values.append(ctypes.c_int(a0_end))
            geometry = plan.geometry(a0_begin, a0_end)
            for extra in getattr(geometry, 'extra', ()):
                values.append(ctypes.c_int(extra))
            pointers = (ctypes.c_void_p * len(values))(
                *[ctypes.cast(ctypes.byref(v), ctypes.c_void_p) for v in values])
            function = plan.function
            if not peer_pointer and getattr(plan, 'function_no_peer', None):
                function = plan.function_no_peer
            grid3 = (ctypes.c_uint32 * 3)(*geometry.grid)
            block3 = (ctypes.c_uint32 * 3)(*geometry.block)
            cluster3 = (ctypes.c_uint32 * 3)(*geometry.cluster)
            prepared = (values, pointers, grid3, block3, cluster3, geometry.smem, function.encode(),
                        _PreparedLaunch(module, function.encode(), grid3, block3, cluster3, geometry.smem, pointers))
            while len(self._launches) >= 1024:
                self._launches.pop(next(iter(self._launches)))
            self._launches[key] = prepared
        if prepare_only:
            return prepared
        # one C call with two arguments: function, geometry and argument block were resolved when the
        # launch was prepared (the per-call cost of the binding is what bounds microsecond stencils)
        code = self._sc_launch_prepared(prepared[7].handle, self.device.stream if stream is None else stream)
        if code != 0:
            check(code, "launch of " + plan.function)

    # -- the call --------------------------------------------------------------------------------
    def __call__(self, *args, **kwargs):
        out = kwargs.pop("out", None)
        iterations = kwargs.pop("iterations", None)
        if kwargs:
            raise StencilException("unexpected keyword arguments {}".format(sorted(kwargs)))
        if iterations is not None:                      # additive: stencil(grid, iterations=n)
            if out is not None:
                raise StencilException("out= and iterations= cannot be combined")
            return self.iterate(args[0], int(iterations), *args[1:])
        self._check_args(args)
        timed = os.environ.get("STENCIL_CUDA_TIMING", "0") not in ("0", "")
        if timed:
            # the reference's generated code reports its own run time through a `duration`
            # out-parameter (backend/c.py:32-39); the analogue here is wall time of the call,
            # device-synchronised, kept on the callable
            import time
            self.device.synchronize()
            begin = time.perf_counter()
        result = self._call_torch(args, out) if _is_torch_tensor(args[0]) else self._call_numpy(args, out)
        if timed:
            self.device.synchronize()
            self.last_duration = time.perf_counter() - begin
        return result

    def _require_device_tensor(self, grid):
        if not grid.is_cuda:
            raise StencilException("torch tensors must live on a CUDA device (numpy arrays are "
                                   "copied for you)")
        if grid.device.index != self.device.index:
            raise StencilException("tensor is on cuda:{} but the stencil was specialised on cuda:{}"
                                   .format(grid.device.index, self.device.index))

    def _check_out_tensor(self, out, tensors):
        """A caller-supplied ``out=`` tensor is written by the kernel as it stands: it must be a
        contiguous CUDA tensor on this device with the grid's shape and dtype, and must not share
        memory with any input (a sweep cannot run in place)."""
        grid = tensors[0]
        if not _is_torch_tensor(out) or not out.is_cuda or out.device.index != self.device.index:
            raise StencilException("out= must be a torch tensor on cuda:{}".format(self.device.index))
        if tuple(out.shape) != tuple(grid.shape) or out.dtype != grid.dtype or not out.is_contiguous():
            raise StencilException("out= must be a contiguous tensor of the grid's shape and dtype")
        first = out.data_ptr()
        last = first + out.numel() * out.element_size()
        for tensor in tensors:
            begin = tensor.data_ptr()
            if begin < last and first < begin + tensor.numel() * tensor.element_size():
                raise StencilException("out= must not overlap an input (a sweep cannot run in place)")

    def _call_torch(self, args, out):
        import torch
        grid = args[0]
        self._require_device_tensor(grid)
        tensors = _device_tensors(args)
        if out is None:
            result = torch.empty_like(tensors[0])
        else:
            self._check_out_tensor(out, tensors)
            result = out
        stream = ctypes.c_void_p(torch.cuda.current_stream(grid.device).cuda_stream)
        self.launch([t.data_ptr() for t in tensors], result.data_ptr(), stream=stream)
        return result

    PIPELINE_MIN_BYTES = 48 << 20      # below this a single copy-in / sweep / copy-out is used
    PIPELINE_CHUNK_BYTES = 32 << 20

    @staticmethod
    def _pipeline_bounds(lo, hi, chunks, min_planes):
        """Axis-0 chunk boundaries of the pipelined host path: `chunks` equal parts, except that
        the first and last ramp up / down (1/8, 1/4, 1/2 of a part) -- nothing overlaps the first
        copy-in and the last copy-out, so they should be short."""
        part = max(1, (hi - lo) // max(1, chunks))
        ramp = [size for size in (part // 8, part // 4, part // 2) if size >= max(1, min_planes)]
        sizes = list(ramp)
        middle = (hi - lo) - 2 * sum(ramp)
        if middle < part:
            sizes, ramp, middle = [], [], hi - lo
        count = max(1, middle // part)
        sizes += [middle // count + (1 if k < middle % count else 0) for k in range(count)]
        sizes += ramp[::-1]
        bounds = [lo]
        for size in sizes:
            bounds.append(bounds[-1] + size)
        assert bounds[-1] == hi
        return bounds

    def _call_numpy(self, args, out):
        """Host arrays in, host array out.  Large grids are pipelined in axis-0 chunks over three
        streams -- copy-in of chunk i+1, sweep of chunk i and copy-out of chunk i-1 overlap, so a
        call costs about one PCIe direction instead of two plus the kernel.  Pinned arrays
        (``pinned_empty``) make the copies truly asynchronous; pageable ones still work."""
        library = load_library()
        device = self.device
        check(library.sc_set_device(device.index))
        arrays = [numpy.ascontiguousarray(a) for a in args]
        grid = arrays[0]
        if out is None:
            result = device.pinned_empty(grid.shape, grid.dtype)
        else:
            if out.shape != grid.shape or out.dtype != grid.dtype or not out.flags.c_contiguous:
                raise StencilException("out= must be a C-contiguous array of the grid's shape and dtype")
            result = out
        buffers = [device.alloc(a.nbytes) for a in arrays]
        out_buffer = device.alloc(grid.nbytes)
        lo, hi = self.plan.domain.sweep_range
        for array, buffer in zip(arrays[1:], buffers[1:]):
            check(library.sc_memcpy_h2d_async(ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(array.ctypes.data),
                                              array.nbytes, device.stream))
        plane_bytes = grid.nbytes // max(1, grid.shape[0])
        self.sweep_host(grid.ctypes.data, (0, grid.shape[0]), [b.ptr for b in buffers], out_buffer.ptr,
                        result.ctypes.data - lo * plane_bytes, (lo, hi),
                        pipelined=grid.ndim >= 2 and self.program.boundary != 'periodic')
        # (periodic: the first chunk reads the last planes -- everything must have landed)
        for buffer in buffers + [out_buffer]:
            buffer.release()
        return result

    def sweep_host(self, host_in, in_range, pointers, out_pointer, host_out, out_range, pipelined=True):
        """One sweep with the grid coming from and the result going to HOST memory, synchronous.

        ``host_in`` is the address of axis-0 planes ``in_range`` = [first, last) of the (local)
        grid; they are copied to the same planes of the device buffer ``pointers[0]`` (tables,
        ``pointers[1:]``, must already be on the device, enqueued on the compute stream).  The
        sweep writes planes ``out_range`` of the device buffer ``out_pointer``, which are copied
        to ``host_out + plane * plane_bytes``.  Large transfers are cut into axis-0 chunks on
        three streams: copy-in of chunk i+1, sweep of chunk i and copy-out of chunk i-1 overlap.
        Used by ``stencil(numpy_array)`` and, per rank, by ``SlabStencil.apply_host``."""
        library = load_library()
        device = self.device
        compute, copy_in, copy_out = device.stream, device.h2d_stream, device.d2h_stream
        shape = self.program.shape
        plane_bytes = self.program.out_dtype_size * int(numpy.prod(shape[1:], dtype=numpy.int64)) \
            if len(shape) > 1 else self.program.out_dtype_size
        in_lo, in_hi = in_range
        lo, hi = out_range
        ghost = self.program.ghost_depth[0]
        nbytes = (hi - lo) * plane_bytes
        chunks = 1
        if pipelined and nbytes >= self.PIPELINE_MIN_BYTES:
            chunk_bytes = int(os.environ.get("STENCIL_CUDA_PIPELINE_CHUNK_MB", "0")) << 20 or self.PIPELINE_CHUNK_BYTES
            chunks = max(1, min(64, nbytes // chunk_bytes, (hi - lo) // max(1, 4 * ghost)))

        def copy_in_planes(begin, end, stream):
            begin, end = max(begin, in_lo), min(end, in_hi)
            if end > begin:
                check(library.sc_memcpy_h2d_async(ctypes.c_void_p(pointers[0] + begin * plane_bytes),
                                                  ctypes.c_void_p(host_in + (begin - in_lo) * plane_bytes),
                                                  (end - begin) * plane_bytes, stream))

        def copy_out_planes(begin, end, stream):
            check(library.sc_memcpy_d2h_async(ctypes.c_void_p(host_out + begin * plane_bytes),
                                              ctypes.c_void_p(out_pointer + begin * plane_bytes),
                                              (end - begin) * plane_bytes, stream))
        if chunks <= 1:
            copy_in_planes(in_lo, in_hi, compute)
            self.launch(pointers, out_pointer, lo, hi, stream=compute)
            copy_out_planes(lo, hi, compute)
            device.synchronize(compute)
            return
        bounds = self._pipeline_bounds(lo, hi, chunks, 4 * ghost)
        chunks = len(bounds) - 1
        # head[k]: the first `ghost` planes of chunk k are on the device -- all that the sweep
        # of chunk k-1 needs from it; landed: everything is
        head = [device.event() for _ in range(chunks)]
        landed = device.event()
        swept = [device.event() for _ in range(chunks)]
        # make sure earlier work on the compute stream (a previous call, the tables) is not overtaken
        fence = device.event()
        check(library.sc_event_record(fence, compute))
        check(library.sc_stream_wait_event(copy_in, fence))
        check(library.sc_stream_wait_event(copy_out, fence))
        for k in range(chunks):
            first = in_lo if k == 0 else bounds[k]                  # planes below the sweep range ride along
            last = in_hi if k == chunks - 1 else bounds[k + 1]
            split = min(bounds[k] + max(ghost, 1), last)
            copy_in_planes(first, split, copy_in)
            check(library.sc_event_record(head[k], copy_in))
            copy_in_planes(split, last, copy_in)
        check(library.sc_event_record(landed, copy_in))
        for k in range(chunks):
            # chunk k reads the first ghost planes of chunk k+1; the copies are in stream order,
            # so those having landed implies that chunk k has
            check(library.sc_stream_wait_event(compute, head[k + 1] if k + 1 < chunks else landed))
            self.launch(pointers, out_pointer, bounds[k], bounds[k + 1], stream=compute)
            check(library.sc_event_record(swept[k], compute))
            check(library.sc_stream_wait_event(copy_out, swept[k]))
            copy_out_planes(bounds[k], bounds[k + 1], copy_out)
        device.synchronize(copy_out)
        device.synchronize(compute)
        for event in head + swept + [fence, landed]:
            device.recycle_event(event)

    # -- temporal blocking ----------------------------------------------------------------------
    def temporal(self):
        """(plan, module) of the fused two-sweep kernel, or None when the stencil is not eligible
        or the grid's planes tile too poorly for it to pay (STENCIL_CUDA_TEMPORAL=0 turns it off,
        =force skips the second test)."""
        if self._temporal is None:
            self._temporal = False
            setting = os.environ.get("STENCIL_CUDA_TEMPORAL", "1")
            if setting not in ("0", ""):
                if self.program.dim == 2:
                    from . import codegen_temporal2d as codegen_temporal      # rows streamed, all in registers
                else:
                    from . import codegen_temporal
                ok, _ = codegen_temporal.eligible(self.program, self.plan.domain)
                if ok and self.plan.strategy == 'stream' and setting != "force":
                    ok = codegen_temporal.worthwhile(
                        self.program, codegen_temporal.choose_config(self.program, self.device.planner_props()))
                if ok and self.plan.strategy == 'stream':
                    plan = codegen_temporal.make_plan(self.program, self.plan.domain,
                                                      self.device.planner_props(), self.plan.options)
                    module = compile_source(plan.source, "stencil_temporal.cu", plan.options)
                    module.load()
                    self._temporal = (plan, module)
        return self._temporal or None

    # -- grids whose rows are no multiple of 16 bytes ---------------------------------------------
    def _pitched_twin(self):
        """The sibling specialisation that works on buffers with rows padded to a multiple of 16
        bytes, or None.  TMA cannot address a grid of, say, 513^3 floats (its row stride must be a
        multiple of 16 bytes), so single calls on such grids take the one-thread-per-point kernel;
        iterate() packs the grid into padded rows once (``sc_memcpy2d_async``), runs all its
        sweeps there at full speed and unpacks the result."""
        if self._twin is None:
            self._twin = False
            program, origin = self.program, getattr(self, 'origin', None)
            width = program.shape[-1]
            from .codegen_stream import _grid_args
            if (origin is not None and width % 4 and program.dim in (2, 3) and program.out_ctype == 'float'
                    and len(_grid_args(program)) == 1 and not self.plan.domain.is_slab
                    and not self.plan.domain.is_pitched
                    and os.environ.get("STENCIL_CUDA_PITCHED", "1") not in ("0", "")
                    and os.environ.get("STENCIL_CUDA_STRATEGY") in (None, "", "stream")):
                try:
                    twin = origin[0].specialize_config(origin[1], row_pitch=-(-width // 4) * 4)
                    if twin.plan.strategy == 'stream':
                        self._twin = twin
                except StencilException:
                    pass
        return self._twin or None

    def _iterate_pitched(self, twin, args, iterations):
        library = load_library()
        device = self.device
        shape = self.program.shape
        width, pitch = shape[-1], twin.plan.domain.row_pitch
        rows = 1
        for extent in shape[:-1]:
            rows *= extent

        def copy2d(dst, dst_pitch, src, src_pitch, stream):
            check(library.sc_memcpy2d_async(ctypes.c_void_p(dst), dst_pitch * 4, ctypes.c_void_p(src),
                                            src_pitch * 4, width * 4, rows, stream), "pitched copy")
        if _is_torch_tensor(args[0]):
            import torch
            stream = ctypes.c_void_p(torch.cuda.current_stream(args[0].device).cuda_stream)
            tensors = _device_tensors(args)
            first = torch.empty(rows * pitch, dtype=tensors[0].dtype, device=tensors[0].device)
            second = torch.empty_like(first)
            copy2d(first.data_ptr(), pitch, tensors[0].data_ptr(), width, stream)
            final = twin.sweeps(first.data_ptr(), second.data_ptr(), [t.data_ptr() for t in tensors[1:]],
                                iterations, stream)
            result = torch.empty_like(tensors[0])
            copy2d(result.data_ptr(), width, final, pitch, stream)
            return result
        check(library.sc_set_device(device.index))
        stream = device.stream
        arrays = [numpy.ascontiguousarray(a) for a in args]
        first, second = device.alloc(rows * pitch * 4), device.alloc(rows * pitch * 4)
        tables = [device.alloc(a.nbytes) for a in arrays[1:]]
        copy2d(first.ptr, pitch, arrays[0].ctypes.data, width, stream)
        for array, buffer in zip(arrays[1:], tables):
            check(library.sc_memcpy_h2d_async(ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(array.ctypes.data),
                                              array.nbytes, stream))
        final = twin.sweeps(first.ptr, second.ptr, [b.ptr for b in tables], iterations, stream)
        result = device.pinned_empty(arrays[0].shape, arrays[0].dtype)
        copy2d(result.ctypes.data, width, final, pitch, stream)
        device.synchronize(stream)
        for buffer in tables + [first, second]:
            buffer.release()
        return result

    def pair_with(self, other):
        """(plan, module) of ONE launch computing ``other(self(grid))`` -- the fused two-sweep kernel
        with a different stencil in each sweep (``fusion.fuse``) -- or None when the pair is not
        eligible, its tiles would not pay, or STENCIL_CUDA_TEMPORAL=0."""
        entry = self._pairs.get(id(other))
        if entry is None:
            if self.program.dim == 2:
                from . import codegen_temporal2d as codegen_temporal
            else:
                from . import codegen_temporal
            found = None
            setting = os.environ.get("STENCIL_CUDA_TEMPORAL", "1")
            if setting not in ("0", "") and self.plan.strategy == 'stream' and other.plan.strategy == 'stream':
                ok, _ = codegen_temporal.pair_eligible(self.program, self.plan.domain,
                                                       other.program, other.plan.domain)
                props = self.device.planner_props()
                if ok and setting != "force":
                    ok = codegen_temporal.worthwhile(
                        self.program, codegen_temporal.choose_config(self.program, props, other.program))
                if ok:
                    plan = codegen_temporal.make_plan(self.program, self.plan.domain, props, self.plan.options,
                                                      second=other.program, second_domain=other.plan.domain)
                    module = compile_source(plan.source, "stencil_pair.cu", plan.options)
                    module.load()
                    found = (plan, module)
            entry = (other, found)                   # keeps `other` alive: its id is the key
            self._pairs[id(other)] = entry
        return entry[1]

    def _dependent_launch_module(self, plan):
        """The plan's kernels compiled with -DSC_PDL (griddepcontrol instructions in), for launches
        with programmatic stream serialization; built on the first iterate()."""
        modules = self._pdl_modules
        module = modules.get(id(plan))
        if module is None:
            module = compile_source(plan.source, "stencil_{}_pdl.cu".format(plan.strategy),
                                    list(plan.options) + ["-DSC_PDL"])
            module.load()
            modules[id(plan)] = module
        return module

    def sweeps(self, pointers_a, pointer_b, table_pointers, count, stream, allow_fused=True):
        """``count`` sweeps ping-ponging between device buffers a and b (a holds the input).
        Pairs of sweeps run as one fused launch where temporal blocking applies.  Returns the
        pointer that holds the result."""
        current, spare = pointers_a, pointer_b
        fused = self.temporal() if allow_fused else None
        stream = self.device.stream if stream is None else stream
        remaining = count

        programmatic = 0 if os.environ.get("STENCIL_CUDA_PDL", "1") in ("0", "") else 1

        def pingpong(plan, module, launches):
            """`launches` launches alternating current->spare, spare->current, enqueued from C"""
            if programmatic:
                module = self._dependent_launch_module(plan)
            even = self._launch_plan(plan, module, [current] + table_pointers, spare, None, None, stream,
                                     prepare_only=True)
            odd = self._launch_plan(plan, module, [spare] + table_pointers, current, None, None, stream,
                                    prepare_only=True)
            if even is None or launches <= 0:
                return
            _, pointers, grid, block, _, smem, name = even[:7]
            check(load_library().sc_launch_pingpong(module.handle, name, grid, block, smem, stream,
                                                    pointers, odd[1], launches, programmatic), "iterated launch")
        if fused is not None and remaining >= 2:
            pingpong(fused[0], fused[1], remaining // 2)
            if (remaining // 2) % 2:
                current, spare = spare, current
            remaining -= 2 * (remaining // 2)
        if remaining > 0:
            pingpong(self.plan, self.module, remaining)
            if remaining % 2:
                current, spare = spare, current
        return current

    # -- iterated application -------------------------------------------------------------------
    def iterate(self, grid, iterations, *tables, **kwargs):
        """``iterations`` sweeps ``g = stencil(g, *tables)`` with the grid resident on the device
        (ping-pong between two buffers); bit-identical to calling the stencil in a loop."""
        if iterations < 0:
            raise StencilException("iterations must be >= 0")
        args = (grid,) + tuple(tables)
        self._check_args(args)
        library = load_library()
        device = self.device
        if _is_torch_tensor(grid):
            self._require_device_tensor(grid)
        twin = self._pitched_twin() if iterations >= 1 else None
        if twin is not None:
            return self._iterate_pitched(twin, args, iterations)
        if _is_torch_tensor(grid):
            import torch
            stream = ctypes.c_void_p(torch.cuda.current_stream(grid.device).cuda_stream)
            tensors = _device_tensors(args)
            if iterations == 0:
                return grid.clone()
            first = torch.empty_like(tensors[0])
            second = torch.empty_like(tensors[0]) if iterations > 1 else None
            tables_ptr = [t.data_ptr() for t in tensors[1:]]
            # the caller's grid is never written: the first (pair of) sweep(s) reads it directly
            fused = self.temporal()
            if fused is not None and iterations >= 2:
                self._launch_plan(fused[0], fused[1], [tensors[0].data_ptr()] + tables_ptr, first.data_ptr(),
                                  None, None, stream)
                done = 2
            else:
                self.launch([tensors[0].data_ptr()] + tables_ptr, first.data_ptr(), stream=stream)
                done = 1
            if done == iterations:
                return first
            if second is None:
                second = torch.empty_like(first)
            result = self.sweeps(first.data_ptr(), second.data_ptr(), tables_ptr, iterations - done, stream)
            return first if result == first.data_ptr() else second
        check(library.sc_set_device(device.index))
        stream = device.stream
        arrays = [numpy.ascontiguousarray(a) for a in args]
        buffers = [device.alloc(a.nbytes) for a in arrays]
        spare = device.alloc(arrays[0].nbytes)
        for array, buffer in zip(arrays, buffers):
            check(library.sc_memcpy_h2d_async(ctypes.c_void_p(buffer.ptr), ctypes.c_void_p(array.ctypes.data),
                                              array.nbytes, stream))
        final = self.sweeps(buffers[0].ptr, spare.ptr, [b.ptr for b in buffers[1:]], iterations, stream)
        current, spare = (buffers[0], spare) if final == buffers[0].ptr else (spare, buffers[0])
        result = device.pinned_empty(arrays[0].shape, arrays[0].dtype)
        check(library.sc_memcpy_d2h_async(ctypes.c_void_p(result.ctypes.data), ctypes.c_void_p(current.ptr),
                                          result.nbytes, stream))
        device.synchronize(stream)
        for buffer in buffers[1:] + [current, spare]:
            buffer.release()
        return result
