// Shared device/host helpers for the seq2seq-vc B200 hot-path kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/s2svc_b200.h"

namespace s2s {

// ---------------------------------------------------------------------------------------------
// error reporting (thread-local last-error string, returned through s2s_last_error())
// ---------------------------------------------------------------------------------------------
char* last_error_buf();
int set_error(int code, const char* fmt, ...);
void count_launch();

#define S2S_REQUIRE(cond, ...)                                         \
    do {                                                               \
        if (!(cond)) return s2s::set_error(S2S_ERR_INVALID, __VA_ARGS__); \
    } while (0)

#define S2S_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return s2s::set_error(S2S_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

#define S2S_LAUNCH_OK()                                                                        \
    do {                                                                                       \
        s2s::count_launch();                                                                   \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return s2s::set_error(S2S_ERR_CUDA, "kernel launch failed: %s (%s:%d)",           \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static inline long ceil_div_l(long a, long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// dtype helpers: activations are f32 or bf16 in HBM, math is always f32 in registers
// ---------------------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-wide vector access (16 B for f32, 8 B for bf16); pointers must be aligned accordingly
template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    }
    static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 t;
        t.x = *reinterpret_cast<uint32_t*>(&a);
        t.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(p) = t;
    }
};

// 8-wide vector access (2 x 16 B for f32, 16 B for bf16); pointers must be 16 B aligned
template <typename T> struct Vec8;
template <> struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Vec8<bf16> {
    static __device__ __forceinline__ void load(const bf16* p, float (&v)[8]) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
    }
    static __device__ __forceinline__ void store(bf16* p, const float (&v)[8]) {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    }
};
static inline bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return al(a) && al(b) && al(c) && al(d);
}

// ---------------------------------------------------------------------------------------------
// Packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): a complex value is one 64-bit register pair, a complex add is one
// instruction and a multiply by a constant twiddle is two (ptxas folds the component swap / sign into operand modifiers).
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float x, float y) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ float2 upk2(u64 r) { float2 d; asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r)); return d; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)));
    return upk2(r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk2(a.x, a.y)), "l"(pk2(b.x, b.y)), "l"(pk2(c.x, c.y)));
    return upk2(r);
}

// runtime dtype -> template dispatch
#define S2S_DISPATCH_DTYPE(dt, T, ...)                         \
    do {                                                       \
        if ((dt) == S2S_F32) { typedef float T; __VA_ARGS__; } \
        else if ((dt) == S2S_BF16) { typedef s2s::bf16 T; __VA_ARGS__; } \
        else return s2s::set_error(S2S_ERR_INVALID, "bad dtype %d", (int)(dt)); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum; `red` must hold >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

// ---------------------------------------------------------------------------------------------
// counter-based dropout RNG: keep(idx) is a pure function of (seed, stream, idx) so the backward
// pass regenerates the forward mask instead of storing it.  p is the DROP probability.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    // murmur3 32-bit finaliser (bijective, full avalanche): 5 integer instructions
    h ^= h >> 16; h *= 0x85ebca6bu;
    h ^= h >> 13; h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}
struct Dropout {
    float p;          // drop probability; 0 disables
    float scale;      // 1/(1-p)
    uint32_t thresh;  // non-zero when dropout is active; a 16-bit lane is dropped iff lane < thresh (= round(p * 2^16))
    uint64_t key;     // mixes seed and stream
    const uint64_t* seed_dev;  // optional device-resident seed increment (CUDA-graph replays)
};
inline Dropout make_dropout(const s2s_dropout_t* d) {
    Dropout r;
    float p = d ? d->p : 0.f;
    r.p = p;
    r.scale = (p > 0.f && p < 1.f) ? 1.f / (1.f - p) : 1.f;
    double t = (double)p * 65536.0 + 0.5;
    r.thresh = (p <= 0.f) ? 0u : (t >= 65535.0 ? 65535u : (t < 1.0 ? 1u : (uint32_t)t));
    uint64_t seed = d ? d->seed : 0, stream_id = d ? d->stream : 0;
    r.key = seed * 0x9e3779b97f4a7c15ull + stream_id * 0xd1b54a32d192ed03ull + 0x2545f4914f6cdd1dull;
    r.seed_dev = d ? d->seed_dev : nullptr;
    return r;
}
// call once per thread at kernel start: folds the device-resident seed into the key
__device__ __forceinline__ void dropout_resolve(Dropout& d) {
    if (d.thresh != 0u && d.seed_dev) d.key += (*d.seed_dev) * 0x9e3779b97f4a7c15ull;
}
// The mask is defined per GROUP of 16 consecutive elements: one murmur-mixed 32-bit seed per group (idx >> 4), and lane
// j = idx & 15 reads the top 16 bits of the j-th state of a 32-bit LCG started at that seed -- x_j = x_0 * A^j + C_j with
// compile-time jump constants, so lane j costs ONE integer multiply-add plus the compare.  (The first version hashed every
// element pair with two murmur rounds: ~10 integer ops per element, which alone doubled the epilogue -- the critical path
// -- of the short-K tcgen05 GEMMs that apply dropout: 16384 x 1536 x 384 took 65 us instead of 33 us.)
__device__ constexpr uint32_t kLcgA[16] = {0x00000001u, 0x915f77f5u, 0xca0bb079u, 0xd21f22cdu, 0x980c9931u, 0xf97362e5u, 0xc6611829u, 0x2d5e2e3du,
                                           0xe8439b61u, 0x4feccad5u, 0x79f220d9u, 0x5b854eadu, 0xbd59b691u, 0xcb881fc5u, 0x6f25fa89u, 0x98a5741du};
__device__ constexpr uint32_t kLcgC[16] = {0x00000000u, 0x3c6ef35fu, 0xc500064au, 0x07d75e31u, 0x8f83df44u, 0x40a83b73u, 0x83bf4d6eu, 0x49542fa5u,
                                           0xaf613f48u, 0x8ca2fb47u, 0x0d906f52u, 0x1cd69ad9u, 0xf753040cu, 0xd23766dbu, 0x62892ff6u, 0x714f33cdu};
// seed of the 16-element group `grp`
__device__ __forceinline__ uint32_t dropout_seed(const Dropout& d, uint64_t grp) {
    uint32_t h = (uint32_t)grp * 0x9e3779b1u + (uint32_t)(grp >> 32) * 0x85ebca77u + (uint32_t)d.key;
    h = mix32(h) ^ (uint32_t)(d.key >> 32);
    h *= 0x2c1b3c6du;
    return h ^ (h >> 15);
}
// LCG state `off` (0..15, runtime) steps after x: binary jump with compile-time constants (predicated multiply-adds, no table)
__device__ __forceinline__ uint32_t dropout_jump(uint32_t x, uint32_t off) {
    x = (off & 8u) ? x * kLcgA[8] + kLcgC[8] : x;
    x = (off & 4u) ? x * kLcgA[4] + kLcgC[4] : x;
    x = (off & 2u) ? x * kLcgA[2] + kLcgC[2] : x;
    x = (off & 1u) ? x * kLcgA[1] + kLcgC[1] : x;
    return x;
}
// returns the multiplicative factor (0 or scale) for element idx
__device__ __forceinline__ float dropout_factor(const Dropout& d, uint64_t idx) {
    if (d.thresh == 0u) return 1.f;
    const uint32_t x = dropout_jump(dropout_seed(d, idx >> 4), (uint32_t)idx & 15u);
    return ((x >> 16) < d.thresh) ? 0.f : d.scale;
}
// factors of N consecutive elements starting at an index that is a multiple of N (N in {2, 4, 8, 16}): one seed, N multiply-adds
template <int N>
__device__ __forceinline__ void dropout_factors(const Dropout& d, uint64_t idx, float (&m)[N]) {
    static_assert(N == 2 || N == 4 || N == 8 || N == 16, "N divides the group of 16");
    if (d.thresh == 0u) {
#pragma unroll
        for (int k = 0; k < N; ++k) m[k] = 1.f;
        return;
    }
    const uint32_t x0 = dropout_jump(dropout_seed(d, idx >> 4), (uint32_t)idx & (15u & ~(uint32_t)(N - 1)));
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const uint32_t x = x0 * kLcgA[k] + kLcgC[k];
        m[k] = ((x >> 16) < d.thresh) ? 0.f : d.scale;
    }
}

// 8 consecutive elements kept as loaded (bf16: one 16-byte register quad) until they are needed as floats
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
    uint4 r;
    __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) { a = reinterpret_cast<const float4*>(p)[0]; b = reinterpret_cast<const float4*>(p)[1]; }
    __device__ __forceinline__ void get(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};


// grid size for grid-stride elementwise kernels: enough CTAs to fill the chip a few times over
inline unsigned ew_grid(long n_items, int per_block) {
    long b = ceil_div_l(n_items, per_block);
    long cap = (long)num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace s2s
