/*
 * stencil_cuda.h -- C ABI of libstencil_cuda.so, the runtime shim of the backend='cuda' path.
 *
 * What it replaces in the reference (paths relative to the reference repository):
 *   - ctree's gcc JIT + on-disk cache behind LazySpecializedFunction / ConcreteStencil._compile
 *     (stencil_code/stencil_kernel.py:61-105, 216-243)            -> sc_compile, sc_module_*
 *   - the ctypes entry point `stencil_kernel(in..., out, float* duration)` built in
 *     SpecializedStencil.finalize (stencil_kernel.py:346-410)      -> sc_launch (+ sc_event_* for
 *     the `duration` the reference measures with clock()/omp_get_wtime(), backend/c.py:32-39)
 *   - the OpenCL host glue: pycl buffer_from_ndarray / buffer_to_ndarray, clEnqueueNDRangeKernel,
 *     clFinish in OclStencilFunction.__call__ and the generated `stencil_control`
 *     (stencil_kernel.py:148-204, backend/ocl.py:163-290)          -> sc_malloc / sc_memcpy_* /
 *     sc_stream_* / sc_launch
 *   - the device query feeding LocalSizeComputer (backend/ocl.py:103-110,
 *     backend/local_size_computer.py:17-40)                        -> sc_device_props
 *   - `int` status + printed message of the OpenCL host code (backend/ocl.py:32-52) and the
 *     StencilException raised on it (stencil_kernel.py:190-196)    -> return codes + sc_last_error
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative SC_E* code, and
 *     leaves a human-readable message for the calling thread in sc_last_error().
 *   - the caller owns every device / host pointer it obtains or passes in.  Pointers allocated
 *     elsewhere in the same process on the same device (e.g. by the PyTorch caching allocator)
 *     are accepted everywhere a device pointer is taken: the shim uses the device's PRIMARY
 *     context, the one the CUDA runtime API uses.
 *   - streams and events are opaque handles (CUstream / CUevent underneath); a NULL stream is the
 *     legacy default stream.  No hidden global stream, no hidden synchronisation.
 *   - libcuda.so.1 and libnvrtc are loaded lazily: the library itself loads on a machine without
 *     a GPU, where sc_compile still works (NVRTC needs no device) and device calls fail with
 *     SC_ENODRIVER.
 *   - thread safety: functions may be called from several threads; the module cache is locked;
 *     a thread that has not called sc_init()/sc_set_device() uses device 0 after sc_init(0).
 */
#ifndef STENCIL_CUDA_H
#define STENCIL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_ABI_VERSION 1

#define SC_OK            0
#define SC_ENODRIVER    -1   /* libcuda.so.1 not loadable or cuInit failed (no GPU)          */
#define SC_ENVRTC       -2   /* libnvrtc not loadable or an NVRTC API error                  */
#define SC_ECOMPILE     -3   /* the kernel source did not compile; see sc_module_log / error */
#define SC_ECUDA        -4   /* a CUDA driver API call failed                                 */
#define SC_EINVAL       -5   /* bad argument                                                  */
#define SC_ENOTLOADED   -6   /* module holds a cubin but is not loaded on a device           */

typedef struct sc_module sc_module;

/* 128-byte, 64-byte aligned TMA descriptor (CUtensorMap). */
typedef struct sc_tmap { uint64_t opaque[16]; } __attribute__((aligned(64))) sc_tmap;

typedef struct sc_device_props {
    char     name[128];
    int32_t  cc_major, cc_minor;
    int32_t  sm_count;
    int32_t  max_threads_per_block;
    int32_t  max_threads_per_sm;
    int32_t  regs_per_sm;
    int32_t  smem_per_sm;            /* bytes */
    int32_t  smem_per_block_optin;   /* bytes */
    int32_t  l2_bytes;
    int32_t  sm_clock_khz, mem_clock_khz;
    int32_t  reserved[4];
    uint64_t total_mem;
} sc_device_props;

/* ---- library / device ------------------------------------------------------------------- */
int         sc_abi_version(void);
const char* sc_last_error(void);
int         sc_driver_available(void);              /* 1 if libcuda loaded and cuInit succeeded */
int         sc_device_count(int* count);
int         sc_init(int device);                    /* retain primary context, make it current  */
int         sc_set_device(int device);              /* make an initialised device current       */
int         sc_get_device(int* device);
int         sc_device_props_get(int device, sc_device_props* out);
int         sc_mem_info(size_t* free_bytes, size_t* total_bytes);
int         sc_device_synchronize(void);

/* ---- JIT: CUDA C++ source -> sm_100a cubin -> loaded module ----------------------------- */
/* Compiles with NVRTC for --gpu-architecture=sm_100a (override: option "--gpu-architecture=...").
 * Identical (source, options) pairs return the cached module.  If a device has been initialised
 * the module is also loaded on it; otherwise it only holds the cubin (sc_module_load later).   */
int         sc_compile(const char* source, const char* name,
                       const char* const* options, int n_options, sc_module** out);
int         sc_module_from_cubin(const void* cubin, size_t size, sc_module** out);
int         sc_module_load(sc_module* module);      /* cuModuleLoadData on the current device   */
int         sc_module_cubin(const sc_module* module, const void** data, size_t* size);
const char* sc_module_log(const sc_module* module); /* NVRTC log ("" if none)                   */
int         sc_module_free(sc_module* module);      /* cached modules are owned by the shim     */
/* attributes of one __global__ function: registers/thread, static smem, local (spill) bytes,
 * max threads per block.                                                                        */
int         sc_function_info(sc_module* module, const char* function,
                             int* regs, int* static_smem, int* local_bytes, int* max_threads);
/* occupancy: resident blocks per SM for a block size and dynamic shared memory size            */
int         sc_function_occupancy(sc_module* module, const char* function,
                                  int block_threads, size_t dynamic_smem, int* blocks_per_sm);

/* ---- TMA descriptor ------------------------------------------------------------------------ */
/* dtype: 0 = float32, 1 = float64, 2 = int32, 3 = uint8.  dims / box are innermost-first
 * (dims[0] is the contiguous axis); strides_bytes has rank-1 entries (dims 1..rank-1).
 * oob_fill: 0 = zeros.  interleave none, swizzle none, L2 promotion 128 B.                     */
int         sc_tensormap_tiled(sc_tmap* out, void* global_address, int dtype, int rank,
                               const uint64_t* dims, const uint64_t* strides_bytes,
                               const uint32_t* box, int oob_fill);

/* ---- launch ---------------------------------------------------------------------------------- */
/* args: array of pointers to the kernel's parameters (cuLaunchKernel convention).  cluster may be
 * NULL or {1,1,1}.  Dynamic shared memory above 48 KB is opted into automatically.              */
int         sc_launch(sc_module* module, const char* function,
                      const uint32_t grid[3], const uint32_t block[3], const uint32_t cluster[3],
                      uint32_t dynamic_smem, void* stream, void** args);

/* `count` launches of one kernel alternating between two argument blocks (iterated sweeps that
 * ping-pong between two buffers: even launches read A and write B, odd ones the reverse), enqueued
 * from C so that short kernels are not bound by the caller's per-launch overhead.  Replaces the
 * Python-level `g = stencil(g)` loop the reference's users write (examples/jacobi_3d.py).
 * See SC_LAUNCH_PROGRAMMATIC below for `flags`.                                                  */
int         sc_launch_pingpong(sc_module* module, const char* function,
                               const uint32_t grid[3], const uint32_t block[3], uint32_t dynamic_smem,
                               void* stream, void** args_even, void** args_odd, int count,
                               uint32_t flags);
/* flags: launch with programmatic stream serialization -- only for kernels that execute
 * griddepcontrol.wait before they touch memory the preceding launch writes or reads            */
#define SC_LAUNCH_PROGRAMMATIC 1u

/* A launch prepared once and repeated: function lookup, shared-memory opt-in and launch
 * configuration are resolved by sc_prepare_launch; sc_launch_prepared is then one call with two
 * arguments (the binding's per-call cost matters for stencils that run in a few microseconds:
 * the reference's 1024 x 1024 timing case).  `args` must stay valid until sc_prepared_free; the
 * parameter values are read at every launch.                                                  */
typedef struct sc_prepared sc_prepared;
int         sc_prepare_launch(sc_module* module, const char* function,
                              const uint32_t grid[3], const uint32_t block[3], const uint32_t cluster[3],
                              uint32_t dynamic_smem, void** args, sc_prepared** prepared);
int         sc_launch_prepared(sc_prepared* prepared, void* stream);
int         sc_prepared_free(sc_prepared* prepared);

/* As sc_launch_pingpong, for kernels that take the number of the step they perform as a 32-bit
 * parameter: launch i receives `first_step + i` in parameter `step_arg` (an index into the args
 * arrays; the pointed-to uint32_t of both blocks is overwritten).  Used by the slab-decomposed
 * iterated sweeps (multi-GPU): every launch is one exchange step whose boundary CTAs wait for /
 * publish the step number in flag words in the neighbouring GPUs' memory, so a whole run of
 * steps is enqueued by one call and the GPUs synchronise among themselves.                      */
int         sc_launch_steps(sc_module* module, const char* function,
                            const uint32_t grid[3], const uint32_t block[3], uint32_t dynamic_smem,
                            void* stream, void** args_even, void** args_odd, int count,
                            uint32_t flags, int step_arg, uint32_t first_step);

/* ---- streams and events --------------------------------------------------------------------- */
int         sc_stream_create(void** stream);
int         sc_stream_destroy(void* stream);
int         sc_stream_synchronize(void* stream);
int         sc_stream_wait_event(void* stream, void* event);
/* stream-ordered flag operations on a 32-bit word in device memory: the stream blocks until
 * *device_ptr >= value (wait), or sets *device_ptr = value once prior work is done (write).  Used
 * with flags that a neighbouring rank's kernel writes through peer memory.                       */
int         sc_stream_wait_value32(void* stream, void* device_ptr, uint32_t value);
int         sc_stream_write_value32(void* stream, void* device_ptr, uint32_t value);
int         sc_event_create(void** event, int with_timing);
int         sc_event_destroy(void* event);
int         sc_event_record(void* event, void* stream);
int         sc_event_synchronize(void* event);
int         sc_event_elapsed_ms(void* start, void* end, float* milliseconds);

/* ---- memory ----------------------------------------------------------------------------------- */
int         sc_malloc(void** device_ptr, size_t bytes);
int         sc_free(void* device_ptr);
int         sc_memset32_async(void* device_ptr, uint32_t value, size_t count, void* stream);
int         sc_memcpy_h2d_async(void* device_dst, const void* host_src, size_t bytes, void* stream);
int         sc_memcpy_d2h_async(void* host_dst, const void* device_src, size_t bytes, void* stream);
int         sc_memcpy_d2d_async(void* device_dst, const void* device_src, size_t bytes, void* stream);
/* `height` rows of `width_bytes`, row r at dst + r*dst_pitch <- src + r*src_pitch; either side host
 * or device memory (unified addressing).  Packs a grid whose rows are no multiple of 16 bytes
 * into the padded row pitch the TMA kernels need, and unpacks the result.                        */
int         sc_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                              size_t width_bytes, size_t height, void* stream);
int         sc_host_alloc(void** host_ptr, size_t bytes);      /* page-locked, portable          */
int         sc_host_free(void* host_ptr);
int         sc_host_register(void* host_ptr, size_t bytes);    /* pin an existing allocation     */
int         sc_host_unregister(void* host_ptr);

/* ---- peer memory between processes on one node (slab halo exchange over NVLink) -------------- */
int         sc_ipc_get_mem_handle(void* device_ptr, unsigned char handle[64]);
int         sc_ipc_open_mem_handle(const unsigned char handle[64], void** device_ptr);
int         sc_ipc_close_mem_handle(void* device_ptr);
int         sc_ipc_get_event_handle(void* event, unsigned char handle[64]);
int         sc_ipc_open_event_handle(const unsigned char handle[64], void** event);
int         sc_event_create_ipc(void** event);                  /* interprocess, no timing        */

#ifdef __cplusplus
}
#endif
#endif /* STENCIL_CUDA_H */
