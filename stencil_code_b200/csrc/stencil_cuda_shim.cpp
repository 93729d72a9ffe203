// libstencil_cuda.so -- C-ABI runtime shim of the backend='cuda' stencil path.
// Interface and the reference call sites it replaces: include/stencil_cuda.h.
//
// Thin on purpose: NVRTC (source -> sm_100a cubin), the CUDA driver API (module load, TMA
// descriptor encode, cuLaunchKernelEx, streams/events/memory, IPC handles) and a module cache.
// libcuda.so.1 and libnvrtc are dlopen'ed on first use so the library loads on a GPU-less host.
#include "stencil_cuda.h"

#include <cuda.h>
#include <nvrtc.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#define SC_STR2(x) #x
#define SC_STR(x) SC_STR2(x)

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buffer[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buffer, sizeof buffer, fmt, ap);
    va_end(ap);
    g_error = buffer;
    return code;
}

// ---- lazily bound driver API ----------------------------------------------------------------
// X(name): `name` is the cuda.h identifier; cuda.h may #define it to a versioned symbol
// (cuMemAlloc -> cuMemAlloc_v2), SC_STR(name) then yields the versioned symbol string.
#define SC_DRIVER_FUNCTIONS(X) \
    X(cuInit) X(cuGetErrorString) X(cuDeviceGet) X(cuDeviceGetCount) X(cuDeviceGetName) \
    X(cuDeviceGetAttribute) X(cuDeviceTotalMem) X(cuDevicePrimaryCtxRetain) X(cuCtxSetCurrent) \
    X(cuCtxGetCurrent) X(cuCtxSynchronize) X(cuMemGetInfo) \
    X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction) X(cuFuncGetAttribute) \
    X(cuFuncSetAttribute) X(cuOccupancyMaxActiveBlocksPerMultiprocessor) X(cuLaunchKernelEx) \
    X(cuTensorMapEncodeTiled) \
    X(cuStreamWaitValue32) X(cuStreamWriteValue32) X(cuStreamCreate) X(cuStreamDestroy) X(cuStreamSynchronize) X(cuStreamWaitEvent) \
    X(cuEventCreate) X(cuEventDestroy) X(cuEventRecord) X(cuEventSynchronize) X(cuEventElapsedTime) \
    X(cuMemAlloc) X(cuMemFree) X(cuMemsetD32Async) X(cuMemcpyHtoDAsync) X(cuMemcpyDtoHAsync) \
    X(cuMemcpyDtoDAsync) X(cuMemcpy2DAsync) X(cuMemHostAlloc) X(cuMemFreeHost) X(cuMemHostRegister) \
    X(cuMemHostUnregister) \
    X(cuIpcGetMemHandle) X(cuIpcOpenMemHandle) X(cuIpcCloseMemHandle) X(cuIpcGetEventHandle) \
    X(cuIpcOpenEventHandle)

struct DriverApi {
#define X(name) decltype(&name) name##_ = nullptr;
    SC_DRIVER_FUNCTIONS(X)
#undef X
    void* handle = nullptr;
    bool tried = false;
    bool ok = false;
    std::string why;
} g_drv;

#define SC_NVRTC_FUNCTIONS(X) \
    X(nvrtcCreateProgram) X(nvrtcDestroyProgram) X(nvrtcCompileProgram) X(nvrtcGetProgramLogSize) \
    X(nvrtcGetProgramLog) X(nvrtcGetCUBINSize) X(nvrtcGetCUBIN) X(nvrtcGetErrorString) X(nvrtcVersion)

struct NvrtcApi {
#define X(name) decltype(&name) name##_ = nullptr;
    SC_NVRTC_FUNCTIONS(X)
#undef X
    void* handle = nullptr;
    bool tried = false;
    bool ok = false;
    std::string why;
} g_rtc;

std::mutex g_mutex;                 // guards lazy binding, device table and module cache

bool bind_driver() {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_drv.tried) return g_drv.ok;
    g_drv.tried = true;
    g_drv.handle = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!g_drv.handle) {
        g_drv.why = std::string("cannot load libcuda.so.1: ") + dlerror();
        return false;
    }
#define X(name)                                                                     \
    g_drv.name##_ = reinterpret_cast<decltype(&name)>(dlsym(g_drv.handle, SC_STR(name))); \
    if (!g_drv.name##_) { g_drv.why = "libcuda.so.1 lacks " SC_STR(name); return false; }
    SC_DRIVER_FUNCTIONS(X)
#undef X
    CUresult rc = g_drv.cuInit_(0);
    if (rc != CUDA_SUCCESS) {
        const char* text = nullptr;
        g_drv.cuGetErrorString_(rc, &text);
        g_drv.why = std::string("cuInit failed: ") + (text ? text : "unknown");
        return false;
    }
    g_drv.ok = true;
    return true;
}

bool bind_nvrtc() {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_rtc.tried) return g_rtc.ok;
    g_rtc.tried = true;
    const char* env = getenv("STENCIL_CUDA_NVRTC");
    const char* candidates[] = {
        env, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
        "/usr/local/cuda/lib64/libnvrtc.so"};
    std::string tried;
    for (const char* path : candidates) {
        if (!path || !*path) continue;
        g_rtc.handle = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        if (g_rtc.handle) break;
        tried += std::string(path) + ": " + dlerror() + "; ";
    }
    if (!g_rtc.handle) {
        g_rtc.why = "cannot load libnvrtc (" + tried + ")";
        return false;
    }
#define X(name)                                                                     \
    g_rtc.name##_ = reinterpret_cast<decltype(&name)>(dlsym(g_rtc.handle, #name));  \
    if (!g_rtc.name##_) { g_rtc.why = "libnvrtc lacks " #name; return false; }
    SC_NVRTC_FUNCTIONS(X)
#undef X
    g_rtc.ok = true;
    return true;
}

int cuda_fail(CUresult rc, const char* what) {
    const char* text = nullptr;
    if (g_drv.cuGetErrorString_) g_drv.cuGetErrorString_(rc, &text);
    return fail(SC_ECUDA, "%s failed: %s (CUresult %d)", what, text ? text : "unknown", (int)rc);
}

#define NEED_DRIVER()                                                              \
    do { if (!bind_driver()) return fail(SC_ENODRIVER, "%s", g_drv.why.c_str()); } while (0)
#define CU(call)                                                                    \
    do { CUresult rc_ = g_drv.call; if (rc_ != CUDA_SUCCESS) return cuda_fail(rc_, #call); } while (0)

// ---- devices -----------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;
CUcontext g_contexts[kMaxDevices] = {};
thread_local int t_device = -1;
int g_default_device = -1;

int ensure_current() {
    NEED_DRIVER();
    int device = t_device >= 0 ? t_device : g_default_device;
    if (device < 0 || !g_contexts[device])
        return fail(SC_EINVAL, "no device initialised; call sc_init(device) first");
    CUcontext current = nullptr;
    CU(cuCtxGetCurrent_(&current));
    if (current != g_contexts[device]) CU(cuCtxSetCurrent_(g_contexts[device]));
    return SC_OK;
}

#define NEED_CONTEXT() do { int rc_ = ensure_current(); if (rc_ != SC_OK) return rc_; } while (0)

}  // namespace

// ---- modules -----------------------------------------------------------------------------------
struct sc_module {
    std::string name;
    std::string log;
    std::vector<char> cubin;
    std::map<int, CUmodule> loaded;              // device -> module
    std::map<std::pair<int, std::string>, CUfunction> functions;
    std::map<std::pair<int, std::string>, int> smem_optin;
    bool cached = false;
};

namespace {
std::map<std::string, sc_module*> g_module_cache;   // key: options + '\0' + source

int current_device_index() { return t_device >= 0 ? t_device : g_default_device; }

int get_function(sc_module* module, const char* function, CUfunction* out) {
    NEED_CONTEXT();
    int device = current_device_index();
    std::lock_guard<std::mutex> lock(g_mutex);
    auto key = std::make_pair(device, std::string(function));
    auto it = module->functions.find(key);
    if (it != module->functions.end()) { *out = it->second; return SC_OK; }
    auto mod = module->loaded.find(device);
    if (mod == module->loaded.end())
        return fail(SC_ENOTLOADED, "module '%s' is not loaded on device %d", module->name.c_str(), device);
    CUfunction fn = nullptr;
    CUresult rc = g_drv.cuModuleGetFunction_(&fn, mod->second, function);
    if (rc != CUDA_SUCCESS) return cuda_fail(rc, (std::string("cuModuleGetFunction(") + function + ")").c_str());
    module->functions[key] = fn;
    *out = fn;
    return SC_OK;
}
}  // namespace

extern "C" {

int sc_abi_version(void) { return SC_ABI_VERSION; }

const char* sc_last_error(void) { return g_error.c_str(); }

int sc_driver_available(void) { return bind_driver() ? 1 : 0; }

int sc_device_count(int* count) {
    if (!count) return fail(SC_EINVAL, "count is NULL");
    NEED_DRIVER();
    CU(cuDeviceGetCount_(count));
    return SC_OK;
}

int sc_init(int device) {
    NEED_DRIVER();
    if (device < 0 || device >= kMaxDevices) return fail(SC_EINVAL, "bad device index %d", device);
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        if (!g_contexts[device]) {
            CUdevice dev;
            CU(cuDeviceGet_(&dev, device));
            CUcontext ctx = nullptr;
            CU(cuDevicePrimaryCtxRetain_(&ctx, dev));
            g_contexts[device] = ctx;
        }
        if (g_default_device < 0) g_default_device = device;
    }
    t_device = device;
    CU(cuCtxSetCurrent_(g_contexts[device]));
    return SC_OK;
}

int sc_set_device(int device) {
    if (device < 0 || device >= kMaxDevices || !g_contexts[device]) return sc_init(device);
    t_device = device;
    return ensure_current();
}

int sc_get_device(int* device) {
    if (!device) return fail(SC_EINVAL, "device is NULL");
    *device = current_device_index();
    return SC_OK;
}

int sc_device_props_get(int device, sc_device_props* out) {
    if (!out) return fail(SC_EINVAL, "out is NULL");
    NEED_DRIVER();
    memset(out, 0, sizeof *out);
    CUdevice dev;
    CU(cuDeviceGet_(&dev, device));
    CU(cuDeviceGetName_(out->name, sizeof out->name, dev));
    size_t total = 0;
    CU(cuDeviceTotalMem_(&total, dev));
    out->total_mem = total;
    struct { int32_t* field; CUdevice_attribute attribute; } table[] = {
        {&out->cc_major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR},
        {&out->cc_minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR},
        {&out->sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT},
        {&out->max_threads_per_block, CU_DEVICE_ATTRIBUTE_MAX_THREADS_PER_BLOCK},
        {&out->max_threads_per_sm, CU_DEVICE_ATTRIBUTE_MAX_THREADS_PER_MULTIPROCESSOR},
        {&out->regs_per_sm, CU_DEVICE_ATTRIBUTE_MAX_REGISTERS_PER_MULTIPROCESSOR},
        {&out->smem_per_sm, CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_MULTIPROCESSOR},
        {&out->smem_per_block_optin, CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN},
        {&out->l2_bytes, CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE},
        {&out->sm_clock_khz, CU_DEVICE_ATTRIBUTE_CLOCK_RATE},
        {&out->mem_clock_khz, CU_DEVICE_ATTRIBUTE_MEMORY_CLOCK_RATE},
    };
    for (auto& row : table) {
        int value = 0;
        CU(cuDeviceGetAttribute_(&value, row.attribute, dev));
        *row.field = value;
    }
    return SC_OK;
}

int sc_mem_info(size_t* free_bytes, size_t* total_bytes) {
    NEED_CONTEXT();
    size_t f = 0, t = 0;
    CU(cuMemGetInfo_(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return SC_OK;
}

int sc_device_synchronize(void) {
    NEED_CONTEXT();
    CU(cuCtxSynchronize_());
    return SC_OK;
}

// ---- JIT ------------------------------------------------------------------------------------
int sc_module_load(sc_module* module) {
    if (!module) return fail(SC_EINVAL, "module is NULL");
    NEED_CONTEXT();
    int device = current_device_index();
    std::lock_guard<std::mutex> lock(g_mutex);
    if (module->loaded.count(device)) return SC_OK;
    if (module->cubin.empty()) return fail(SC_EINVAL, "module '%s' holds no cubin", module->name.c_str());
    CUmodule mod = nullptr;
    CUresult rc = g_drv.cuModuleLoadData_(&mod, module->cubin.data());
    if (rc != CUDA_SUCCESS) return cuda_fail(rc, "cuModuleLoadData");
    module->loaded[device] = mod;
    return SC_OK;
}

int sc_compile(const char* source, const char* name, const char* const* options, int n_options,
               sc_module** out) {
    if (!source || !out) return fail(SC_EINVAL, "source/out is NULL");
    if (!bind_nvrtc()) return fail(SC_ENVRTC, "%s", g_rtc.why.c_str());
    std::vector<std::string> opts;
    bool has_arch = false;
    for (int i = 0; i < n_options; ++i) {
        opts.emplace_back(options[i]);
        if (opts.back().rfind("--gpu-architecture", 0) == 0 || opts.back().rfind("-arch", 0) == 0)
            has_arch = true;
    }
    if (!has_arch) opts.emplace_back("--gpu-architecture=sm_100a");
    std::string key;
    for (auto& o : opts) { key += o; key.push_back('\n'); }
    key.push_back('\0');
    key += source;

    sc_module* module = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_module_cache.find(key);
        if (it != g_module_cache.end()) module = it->second;
    }
    if (!module) {
        nvrtcProgram program = nullptr;
        const char* program_name = name ? name : "stencil_kernel.cu";
        nvrtcResult rc = g_rtc.nvrtcCreateProgram_(&program, source, program_name, 0, nullptr, nullptr);
        if (rc != NVRTC_SUCCESS)
            return fail(SC_ENVRTC, "nvrtcCreateProgram: %s", g_rtc.nvrtcGetErrorString_(rc));
        std::vector<const char*> raw;
        for (auto& o : opts) raw.push_back(o.c_str());
        rc = g_rtc.nvrtcCompileProgram_(program, (int)raw.size(), raw.data());
        std::string log;
        size_t log_size = 0;
        if (g_rtc.nvrtcGetProgramLogSize_(program, &log_size) == NVRTC_SUCCESS && log_size > 1) {
            log.resize(log_size);
            g_rtc.nvrtcGetProgramLog_(program, &log[0]);
            while (!log.empty() && log.back() == '\0') log.pop_back();
        }
        if (rc != NVRTC_SUCCESS) {
            g_rtc.nvrtcDestroyProgram_(&program);
            return fail(SC_ECOMPILE, "NVRTC could not compile '%s': %s\n%s", program_name,
                        g_rtc.nvrtcGetErrorString_(rc), log.c_str());
        }
        size_t cubin_size = 0;
        rc = g_rtc.nvrtcGetCUBINSize_(program, &cubin_size);
        if (rc != NVRTC_SUCCESS || cubin_size == 0) {
            g_rtc.nvrtcDestroyProgram_(&program);
            return fail(SC_ENVRTC, "nvrtcGetCUBINSize: %s (a real sm_ architecture is required)",
                        g_rtc.nvrtcGetErrorString_(rc));
        }
        module = new sc_module();
        module->name = program_name;
        module->log = log;
        module->cubin.resize(cubin_size);
        g_rtc.nvrtcGetCUBIN_(program, module->cubin.data());
        g_rtc.nvrtcDestroyProgram_(&program);
        module->cached = true;
        std::lock_guard<std::mutex> lock(g_mutex);
        auto inserted = g_module_cache.emplace(key, module);
        if (!inserted.second) { delete module; module = inserted.first->second; }
    }
    *out = module;
    // load on the current device if there is one; a GPU-less host keeps the cubin only
    if (g_drv.ok && current_device_index() >= 0) return sc_module_load(module);
    return SC_OK;
}

int sc_module_from_cubin(const void* cubin, size_t size, sc_module** out) {
    if (!cubin || !size || !out) return fail(SC_EINVAL, "cubin/out is NULL or empty");
    sc_module* module = new sc_module();
    module->name = "cubin";
    module->cubin.assign((const char*)cubin, (const char*)cubin + size);
    *out = module;
    if (g_drv.ok && current_device_index() >= 0) return sc_module_load(module);
    return SC_OK;
}

int sc_module_cubin(const sc_module* module, const void** data, size_t* size) {
    if (!module || !data || !size) return fail(SC_EINVAL, "NULL argument");
    *data = module->cubin.data();
    *size = module->cubin.size();
    return SC_OK;
}

const char* sc_module_log(const sc_module* module) { return module ? module->log.c_str() : ""; }

int sc_module_free(sc_module* module) {
    if (!module || module->cached) return SC_OK;     // cached modules live as long as the library
    if (g_drv.ok) {
        for (auto& entry : module->loaded) {
            if (g_contexts[entry.first]) {
                g_drv.cuCtxSetCurrent_(g_contexts[entry.first]);
                g_drv.cuModuleUnload_(entry.second);
            }
        }
        if (current_device_index() >= 0) ensure_current();
    }
    delete module;
    return SC_OK;
}

int sc_function_info(sc_module* module, const char* function, int* regs, int* static_smem,
                     int* local_bytes, int* max_threads) {
    if (!module || !function) return fail(SC_EINVAL, "NULL argument");
    CUfunction fn;
    int rc = get_function(module, function, &fn);
    if (rc != SC_OK) return rc;
    if (regs) CU(cuFuncGetAttribute_(regs, CU_FUNC_ATTRIBUTE_NUM_REGS, fn));
    if (static_smem) CU(cuFuncGetAttribute_(static_smem, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, fn));
    if (local_bytes) CU(cuFuncGetAttribute_(local_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, fn));
    if (max_threads) CU(cuFuncGetAttribute_(max_threads, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, fn));
    return SC_OK;
}

static int optin_smem(sc_module* module, const char* function, CUfunction fn, size_t dynamic_smem) {
    // static + dynamic above 48 KB needs the opt-in too, so any non-trivial dynamic size is declared
    if (dynamic_smem <= 8 * 1024) return SC_OK;
    auto key = std::make_pair(current_device_index(), std::string(function));
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = module->smem_optin.find(key);
        if (it != module->smem_optin.end() && it->second >= (int)dynamic_smem) return SC_OK;
    }
    CU(cuFuncSetAttribute_(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dynamic_smem));
    std::lock_guard<std::mutex> lock(g_mutex);
    module->smem_optin[key] = (int)dynamic_smem;
    return SC_OK;
}

int sc_function_occupancy(sc_module* module, const char* function, int block_threads,
                          size_t dynamic_smem, int* blocks_per_sm) {
    if (!module || !function || !blocks_per_sm) return fail(SC_EINVAL, "NULL argument");
    CUfunction fn;
    int rc = get_function(module, function, &fn);
    if (rc != SC_OK) return rc;
    rc = optin_smem(module, function, fn, dynamic_smem);
    if (rc != SC_OK) return rc;
    CU(cuOccupancyMaxActiveBlocksPerMultiprocessor_(blocks_per_sm, fn, block_threads, dynamic_smem));
    return SC_OK;
}

// ---- TMA descriptor ----------------------------------------------------------------------------
int sc_tensormap_tiled(sc_tmap* out, void* global_address, int dtype, int rank,
                       const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                       int oob_fill) {
    if (!out || !global_address || !dims || !box) return fail(SC_EINVAL, "NULL argument");
    if (rank < 1 || rank > 5) return fail(SC_EINVAL, "tensor map rank %d outside 1..5", rank);
    NEED_CONTEXT();
    static_assert(sizeof(sc_tmap) == sizeof(CUtensorMap), "sc_tmap must match CUtensorMap");
    CUtensorMapDataType type;
    switch (dtype) {
        case 0: type = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
        case 1: type = CU_TENSOR_MAP_DATA_TYPE_FLOAT64; break;
        case 2: type = CU_TENSOR_MAP_DATA_TYPE_INT32; break;
        case 3: type = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
        default: return fail(SC_EINVAL, "unknown tensor map dtype %d", dtype);
    }
    cuuint64_t global_dims[5];
    cuuint64_t global_strides[5] = {0, 0, 0, 0, 0};
    cuuint32_t box_dims[5];
    cuuint32_t element_strides[5];
    for (int i = 0; i < rank; ++i) {
        global_dims[i] = dims[i];
        box_dims[i] = box[i];
        element_strides[i] = 1;
        if (i + 1 < rank) {
            if (!strides_bytes) return fail(SC_EINVAL, "strides_bytes is NULL for rank %d", rank);
            global_strides[i] = strides_bytes[i];
        }
    }
    // L2 promotion: 128 B by default; STENCIL_CUDA_L2_PROMOTION=0|64|128|256 overrides (tuning knob)
    CUtensorMapL2promotion promotion = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (const char* env = getenv("STENCIL_CUDA_L2_PROMOTION")) {
        switch (atoi(env)) {
            case 0: promotion = CU_TENSOR_MAP_L2_PROMOTION_NONE; break;
            case 64: promotion = CU_TENSOR_MAP_L2_PROMOTION_L2_64B; break;
            case 256: promotion = CU_TENSOR_MAP_L2_PROMOTION_L2_256B; break;
            default: promotion = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; break;
        }
    }
    CUresult rc = g_drv.cuTensorMapEncodeTiled_(
        reinterpret_cast<CUtensorMap*>(out), type, (cuuint32_t)rank, global_address, global_dims,
        global_strides, box_dims, element_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, promotion,
        oob_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return cuda_fail(rc, "cuTensorMapEncodeTiled");
    return SC_OK;
}

// ---- launch --------------------------------------------------------------------------------------
int sc_launch(sc_module* module, const char* function, const uint32_t grid[3],
              const uint32_t block[3], const uint32_t cluster[3], uint32_t dynamic_smem,
              void* stream, void** args) {
    if (!module || !function || !grid || !block) return fail(SC_EINVAL, "NULL argument");
    CUfunction fn;
    int rc = get_function(module, function, &fn);
    if (rc != SC_OK) return rc;
    rc = optin_smem(module, function, fn, dynamic_smem);
    if (rc != SC_OK) return rc;
    CUlaunchConfig config;
    memset(&config, 0, sizeof config);
    config.gridDimX = grid[0]; config.gridDimY = grid[1]; config.gridDimZ = grid[2];
    config.blockDimX = block[0]; config.blockDimY = block[1]; config.blockDimZ = block[2];
    config.sharedMemBytes = dynamic_smem;
    config.hStream = reinterpret_cast<CUstream>(stream);
    CUlaunchAttribute attribute;
    if (cluster && (cluster[0] > 1 || cluster[1] > 1 || cluster[2] > 1)) {
        memset(&attribute, 0, sizeof attribute);
        attribute.id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
        attribute.value.clusterDim.x = cluster[0];
        attribute.value.clusterDim.y = cluster[1];
        attribute.value.clusterDim.z = cluster[2];
        config.attrs = &attribute;
        config.numAttrs = 1;
    }
    CUresult cu = g_drv.cuLaunchKernelEx_(&config, fn, args, nullptr);
    if (cu != CUDA_SUCCESS) return cuda_fail(cu, (std::string("cuLaunchKernelEx(") + function + ")").c_str());
    return SC_OK;
}

int sc_launch_pingpong(sc_module* module, const char* function, const uint32_t grid[3],
                       const uint32_t block[3], uint32_t dynamic_smem, void* stream,
                       void** args_even, void** args_odd, int count, uint32_t flags) {
    if (!module || !function || !grid || !block || !args_even || !args_odd)
        return fail(SC_EINVAL, "NULL argument");
    CUfunction fn;
    int rc = get_function(module, function, &fn);
    if (rc != SC_OK) return rc;
    rc = optin_smem(module, function, fn, dynamic_smem);
    if (rc != SC_OK) return rc;
    CUlaunchConfig config;
    memset(&config, 0, sizeof config);
    config.gridDimX = grid[0]; config.gridDimY = grid[1]; config.gridDimZ = grid[2];
    config.blockDimX = block[0]; config.blockDimY = block[1]; config.blockDimZ = block[2];
    config.sharedMemBytes = dynamic_smem;
    config.hStream = reinterpret_cast<CUstream>(stream);
    CUlaunchAttribute attribute;
    memset(&attribute, 0, sizeof attribute);
    if (flags & SC_LAUNCH_PROGRAMMATIC) {
        attribute.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
        attribute.value.programmaticStreamSerializationAllowed = 1;
        config.attrs = &attribute;
        config.numAttrs = 1;
    }
    for (int i = 0; i < count; ++i) {
        CUresult cu = g_drv.cuLaunchKernelEx_(&config, fn, (i & 1) ? args_odd : args_even, nullptr);
        if (cu != CUDA_SUCCESS)
            return cuda_fail(cu, (std::string("cuLaunchKernelEx(") + function + ")").c_str());
    }
    return SC_OK;
}

struct sc_prepared {
    CUfunction fn;
    CUlaunchConfig config;
    CUlaunchAttribute attribute;
    void** args;
    int device;
};

int sc_prepare_launch(sc_module* module, const char* function, const uint32_t grid[3],
                      const uint32_t block[3], const uint32_t cluster[3], uint32_t dynamic_smem,
                      void** args, sc_prepared** prepared) {
    if (!module || !function || !grid || !block || !prepared) return fail(SC_EINVAL, "NULL argument");
    CUfunction fn;
    int rc = get_function(module, function, &fn);
    if (rc != SC_OK) return rc;
    rc = optin_smem(module, function, fn, dynamic_smem);
    if (rc != SC_OK) return rc;
    sc_prepared* p = new sc_prepared();
    p->fn = fn;
    p->args = args;
    p->device = current_device_index();
    memset(&p->config, 0, sizeof p->config);
    p->config.gridDimX = grid[0]; p->config.gridDimY = grid[1]; p->config.gridDimZ = grid[2];
    p->config.blockDimX = block[0]; p->config.blockDimY = block[1]; p->config.blockDimZ = block[2];
    p->config.sharedMemBytes = dynamic_smem;
    memset(&p->attribute, 0, sizeof p->attribute);
    if (cluster && (cluster[0] > 1 || cluster[1] > 1 || cluster[2] > 1)) {
        p->attribute.id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
        p->attribute.value.clusterDim.x = cluster[0];
        p->attribute.value.clusterDim.y = cluster[1];
        p->attribute.value.clusterDim.z = cluster[2];
        p->config.attrs = &p->attribute;
        p->config.numAttrs = 1;
    }
    *prepared = p;
    return SC_OK;
}

int sc_launch_prepared(sc_prepared* prepared, void* stream) {
    if (!prepared) return fail(SC_EINVAL, "prepared is NULL");
    CUlaunchConfig config = prepared->config;           // a copy: the same prepared launch may be used on several streams
    config.hStream = reinterpret_cast<CUstream>(stream);
    CUresult cu = g_drv.cuLaunchKernelEx_(&config, prepared->fn, prepared->args, nullptr);
    if (cu != CUDA_SUCCESS) return cuda_fail(cu, "cuLaunchKernelEx(prepared)");
    return SC_OK;
}

int sc_prepared_free(sc_prepared* prepared) {
    delete prepared;
    return SC_OK;
}

int sc_launch_steps(sc_module* module, const char* function, const uint32_t grid[3],
                    const uint32_t block[3], uint32_t dynamic_smem, void* stream,
                    void** args_even, void** args_odd, int count, uint32_t flags,
                    int step_arg, uint32_t first_step) {
    if (!module || !function || !grid || !block || !args_even || !args_odd)
        return fail(SC_EINVAL, "NULL argument");
    if (step_arg < 0 || !args_even[step_arg] || !args_odd[step_arg])
        return fail(SC_EINVAL, "step_arg does not name a parameter");
    CUfunction fn;
    int rc = get_function(module, function, &fn);
    if (rc != SC_OK) return rc;
    rc = optin_smem(module, function, fn, dynamic_smem);
    if (rc != SC_OK) return rc;
    CUlaunchConfig config;
    memset(&config, 0, sizeof config);
    config.gridDimX = grid[0]; config.gridDimY = grid[1]; config.gridDimZ = grid[2];
    config.blockDimX = block[0]; config.blockDimY = block[1]; config.blockDimZ = block[2];
    config.sharedMemBytes = dynamic_smem;
    config.hStream = reinterpret_cast<CUstream>(stream);
    CUlaunchAttribute attribute;
    memset(&attribute, 0, sizeof attribute);
    if (flags & SC_LAUNCH_PROGRAMMATIC) {
        attribute.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
        attribute.value.programmaticStreamSerializationAllowed = 1;
        config.attrs = &attribute;
        config.numAttrs = 1;
    }
    for (int i = 0; i < count; ++i) {
        void** args = (i & 1) ? args_odd : args_even;
        *static_cast<uint32_t*>(args[step_arg]) = first_step + (uint32_t)i;   // copied by the launch
        CUresult cu = g_drv.cuLaunchKernelEx_(&config, fn, args, nullptr);
        if (cu != CUDA_SUCCESS)
            return cuda_fail(cu, (std::string("cuLaunchKernelEx(") + function + ")").c_str());
    }
    return SC_OK;
}

// ---- streams and events ------------------------------------------------------------------------
int sc_stream_create(void** stream) {
    if (!stream) return fail(SC_EINVAL, "stream is NULL");
    NEED_CONTEXT();
    CUstream s = nullptr;
    CU(cuStreamCreate_(&s, CU_STREAM_NON_BLOCKING));
    *stream = s;
    return SC_OK;
}

int sc_stream_destroy(void* stream) {
    NEED_CONTEXT();
    CU(cuStreamDestroy_(reinterpret_cast<CUstream>(stream)));
    return SC_OK;
}

int sc_stream_synchronize(void* stream) {
    NEED_CONTEXT();
    CU(cuStreamSynchronize_(reinterpret_cast<CUstream>(stream)));
    return SC_OK;
}

int sc_stream_wait_event(void* stream, void* event) {
    NEED_CONTEXT();
    CU(cuStreamWaitEvent_(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUevent>(event), 0));
    return SC_OK;
}

int sc_stream_wait_value32(void* stream, void* device_ptr, uint32_t value) {
    NEED_CONTEXT();
    CU(cuStreamWaitValue32_(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(device_ptr),
                            value, CU_STREAM_WAIT_VALUE_GEQ));
    return SC_OK;
}

int sc_stream_write_value32(void* stream, void* device_ptr, uint32_t value) {
    NEED_CONTEXT();
    CU(cuStreamWriteValue32_(reinterpret_cast<CUstream>(stream), reinterpret_cast<CUdeviceptr>(device_ptr),
                             value, CU_STREAM_WRITE_VALUE_DEFAULT));
    return SC_OK;
}

int sc_event_create(void** event, int with_timing) {
    if (!event) return fail(SC_EINVAL, "event is NULL");
    NEED_CONTEXT();
    CUevent e = nullptr;
    CU(cuEventCreate_(&e, with_timing ? CU_EVENT_DEFAULT : CU_EVENT_DISABLE_TIMING));
    *event = e;
    return SC_OK;
}

int sc_event_create_ipc(void** event) {
    if (!event) return fail(SC_EINVAL, "event is NULL");
    NEED_CONTEXT();
    CUevent e = nullptr;
    CU(cuEventCreate_(&e, CU_EVENT_DISABLE_TIMING | CU_EVENT_INTERPROCESS));
    *event = e;
    return SC_OK;
}

int sc_event_destroy(void* event) {
    NEED_CONTEXT();
    CU(cuEventDestroy_(reinterpret_cast<CUevent>(event)));
    return SC_OK;
}

int sc_event_record(void* event, void* stream) {
    NEED_CONTEXT();
    CU(cuEventRecord_(reinterpret_cast<CUevent>(event), reinterpret_cast<CUstream>(stream)));
    return SC_OK;
}

int sc_event_synchronize(void* event) {
    NEED_CONTEXT();
    CU(cuEventSynchronize_(reinterpret_cast<CUevent>(event)));
    return SC_OK;
}

int sc_event_elapsed_ms(void* start, void* end, float* milliseconds) {
    if (!milliseconds) return fail(SC_EINVAL, "milliseconds is NULL");
    NEED_CONTEXT();
    CU(cuEventElapsedTime_(milliseconds, reinterpret_cast<CUevent>(start), reinterpret_cast<CUevent>(end)));
    return SC_OK;
}

// ---- memory -----------------------------------------------------------------------------------
int sc_malloc(void** device_ptr, size_t bytes) {
    if (!device_ptr) return fail(SC_EINVAL, "device_ptr is NULL");
    NEED_CONTEXT();
    CUdeviceptr p = 0;
    CU(cuMemAlloc_(&p, bytes ? bytes : 1));
    *device_ptr = reinterpret_cast<void*>(p);
    return SC_OK;
}

int sc_free(void* device_ptr) {
    if (!device_ptr) return SC_OK;
    NEED_CONTEXT();
    CU(cuMemFree_(reinterpret_cast<CUdeviceptr>(device_ptr)));
    return SC_OK;
}

int sc_memset32_async(void* device_ptr, uint32_t value, size_t count, void* stream) {
    NEED_CONTEXT();
    CU(cuMemsetD32Async_(reinterpret_cast<CUdeviceptr>(device_ptr), value, count,
                         reinterpret_cast<CUstream>(stream)));
    return SC_OK;
}

int sc_memcpy_h2d_async(void* device_dst, const void* host_src, size_t bytes, void* stream) {
    NEED_CONTEXT();
    CU(cuMemcpyHtoDAsync_(reinterpret_cast<CUdeviceptr>(device_dst), host_src, bytes,
                          reinterpret_cast<CUstream>(stream)));
    return SC_OK;
}

int sc_memcpy_d2h_async(void* host_dst, const void* device_src, size_t bytes, void* stream) {
    NEED_CONTEXT();
    CU(cuMemcpyDtoHAsync_(host_dst, reinterpret_cast<CUdeviceptr>(device_src), bytes,
                          reinterpret_cast<CUstream>(stream)));
    return SC_OK;
}

int sc_memcpy_d2d_async(void* device_dst, const void* device_src, size_t bytes, void* stream) {
    NEED_CONTEXT();
    CU(cuMemcpyDtoDAsync_(reinterpret_cast<CUdeviceptr>(device_dst),
                          reinterpret_cast<CUdeviceptr>(device_src), bytes,
                          reinterpret_cast<CUstream>(stream)));
    return SC_OK;
}

int sc_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                      size_t width_bytes, size_t height, void* stream) {
    if (!dst || !src) return fail(SC_EINVAL, "NULL pointer");
    if (width_bytes > dst_pitch || width_bytes > src_pitch) return fail(SC_EINVAL, "row wider than its pitch");
    NEED_CONTEXT();
    CUDA_MEMCPY2D copy;
    memset(&copy, 0, sizeof copy);
    copy.srcMemoryType = CU_MEMORYTYPE_UNIFIED;          // host or device: resolved by address
    copy.srcDevice = reinterpret_cast<CUdeviceptr>(src);
    copy.srcPitch = src_pitch;
    copy.dstMemoryType = CU_MEMORYTYPE_UNIFIED;
    copy.dstDevice = reinterpret_cast<CUdeviceptr>(dst);
    copy.dstPitch = dst_pitch;
    copy.WidthInBytes = width_bytes;
    copy.Height = height;
    if (width_bytes && height)
        CU(cuMemcpy2DAsync_(&copy, reinterpret_cast<CUstream>(stream)));
    return SC_OK;
}

int sc_host_alloc(void** host_ptr, size_t bytes) {
    if (!host_ptr) return fail(SC_EINVAL, "host_ptr is NULL");
    NEED_CONTEXT();
    CU(cuMemHostAlloc_(host_ptr, bytes ? bytes : 1, CU_MEMHOSTALLOC_PORTABLE));
    return SC_OK;
}

int sc_host_free(void* host_ptr) {
    if (!host_ptr) return SC_OK;
    NEED_CONTEXT();
    CU(cuMemFreeHost_(host_ptr));
    return SC_OK;
}

int sc_host_register(void* host_ptr, size_t bytes) {
    NEED_CONTEXT();
    CU(cuMemHostRegister_(host_ptr, bytes, CU_MEMHOSTREGISTER_PORTABLE));
    return SC_OK;
}

int sc_host_unregister(void* host_ptr) {
    NEED_CONTEXT();
    CU(cuMemHostUnregister_(host_ptr));
    return SC_OK;
}

// ---- IPC -----------------------------------------------------------------------------------------
int sc_ipc_get_mem_handle(void* device_ptr, unsigned char handle[64]) {
    NEED_CONTEXT();
    static_assert(sizeof(CUipcMemHandle) == 64, "CUipcMemHandle is 64 bytes");
    CUipcMemHandle h;
    CU(cuIpcGetMemHandle_(&h, reinterpret_cast<CUdeviceptr>(device_ptr)));
    memcpy(handle, &h, 64);
    return SC_OK;
}

int sc_ipc_open_mem_handle(const unsigned char handle[64], void** device_ptr) {
    if (!device_ptr) return fail(SC_EINVAL, "device_ptr is NULL");
    NEED_CONTEXT();
    CUipcMemHandle h;
    memcpy(&h, handle, 64);
    CUdeviceptr p = 0;
    CU(cuIpcOpenMemHandle_(&p, h, CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS));
    *device_ptr = reinterpret_cast<void*>(p);
    return SC_OK;
}

int sc_ipc_close_mem_handle(void* device_ptr) {
    NEED_CONTEXT();
    CU(cuIpcCloseMemHandle_(reinterpret_cast<CUdeviceptr>(device_ptr)));
    return SC_OK;
}

int sc_ipc_get_event_handle(void* event, unsigned char handle[64]) {
    NEED_CONTEXT();
    static_assert(sizeof(CUipcEventHandle) == 64, "CUipcEventHandle is 64 bytes");
    CUipcEventHandle h;
    CU(cuIpcGetEventHandle_(&h, reinterpret_cast<CUevent>(event)));
    memcpy(handle, &h, 64);
    return SC_OK;
}

int sc_ipc_open_event_handle(const unsigned char handle[64], void** event) {
    if (!event) return fail(SC_EINVAL, "event is NULL");
    NEED_CONTEXT();
    CUipcEventHandle h;
    memcpy(&h, handle, 64);
    CUevent e = nullptr;
    CU(cuIpcOpenEventHandle_(&e, h));
    *event = e;
    return SC_OK;
}

}  // extern "C"
