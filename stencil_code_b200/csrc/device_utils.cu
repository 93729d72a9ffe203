// Fixed helper kernels of the stencil CUDA backend (everything stencil-specific is generated at
// run time by backend/cuda/*; these few are shape-agnostic utilities).  Built to a cubin by
// csrc/Makefile (nvcc, sm_100a) and loaded through sc_module_from_cubin; if the cubin is absent
// the same file is compiled through NVRTC.
//
//   sc_fill_uniform   counter-based uniform [0, scale) float32 fill -- synthetic benchmark grids
//                     that are too large to generate on the host
//   sc_checksum       order-independent 64-bit checksum of a float32 grid (bit patterns), used to
//                     compare 1-GPU and k-GPU results at sizes no CPU oracle can reach
//   sc_count_diff     number of elements whose bit patterns differ between two grids
//   sc_signal         publish "my boundary planes for sweep `value` are in your ghost planes":
//                     system-scope fence, then a flag store into the neighbour's (peer-mapped) memory
typedef unsigned int u32;
typedef unsigned long long u64;

__device__ __forceinline__ u32 sc_mix(u64 x) {
    // splitmix64 finaliser; the high 24 bits feed the mantissa
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x = x ^ (x >> 31);
    return (u32)(x >> 40);
}

extern "C" __global__ void __launch_bounds__(256)
sc_fill_uniform(float* __restrict__ out, u64 count, u64 seed, u64 first_index, float scale) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const u32 bits = sc_mix((first_index + i) ^ (seed * 0xD1342543DE82EF95ull));
        out[i] = (float)bits * (1.0f / 16777216.0f) * scale;
    }
}

extern "C" __global__ void __launch_bounds__(256)
sc_checksum(const u32* __restrict__ in, u64 count, u64 first_index, u64* __restrict__ result) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u64 local = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const u64 position = first_index + i;
        u64 h = ((u64)in[i] + 0x9E3779B97F4A7C15ull) * (2 * position + 1);
        h ^= h >> 29;
        local += h * 0xBF58476D1CE4E5B9ull;
    }
    for (int offset = 16; offset > 0; offset >>= 1)
        local += __shfl_xor_sync(0xffffffffu, local, offset);
    if ((threadIdx.x & 31) == 0) atomicAdd(result, local);
}

extern "C" __global__ void __launch_bounds__(256)
sc_count_diff(const u32* __restrict__ a, const u32* __restrict__ b, u64 count, u64* __restrict__ result) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u64 local = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        local += (a[i] != b[i]) ? 1 : 0;
    for (int offset = 16; offset > 0; offset >>= 1)
        local += __shfl_xor_sync(0xffffffffu, local, offset);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(result, local);
}

extern "C" __global__ void sc_signal(volatile u32* flag_a, volatile u32* flag_b, u32 value) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence_system();
        if (flag_a) *flag_a = value;
        if (flag_b) *flag_b = value;
        __threadfence_system();
    }
}
