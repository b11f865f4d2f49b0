// HBM-bound normalisation kernels: LayerNorm fwd/bwd, column reductions (bias / affine grads),
// relu / dropout backward, and the postnet BatchNorm(+tanh) pipeline over zero-haloed
// channels-last buffers.  All kernels read each input once per pass with 4-wide vector access
// when the channel count allows it; statistics and parameter gradients are float32.
#include "common.cuh"

namespace s2s {

template <typename T, int VEC> struct VLoad;
template <typename T> struct VLoad<T, 4> {
    static __device__ __forceinline__ void ld(const T* p, float (&v)[4]) { Vec4<T>::load(p, v); }
    static __device__ __forceinline__ void st(T* p, const float (&v)[4]) { Vec4<T>::store(p, v); }
};
template <typename T> struct VLoad<T, 8> {
    static __device__ __forceinline__ void ld(const T* p, float (&v)[8]) { Vec8<T>::load(p, v); }
    static __device__ __forceinline__ void st(T* p, const float (&v)[8]) { Vec8<T>::store(p, v); }
};
template <typename T> struct VLoad<T, 1> {
    static __device__ __forceinline__ void ld(const T* p, float (&v)[1]) { v[0] = to_f<T>(*p); }
    static __device__ __forceinline__ void st(T* p, const float (&v)[1]) { *p = from_f<T>(v[0]); }
};

static inline bool vec4_ok(int cols, int64_t ld, const void* a, const void* b = nullptr, const void* c = nullptr,
                           const void* d = nullptr) {
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return (cols % 4 == 0) && (ld % 4 == 0) && al(a) && al(b) && al(c) && al(d);
}

// =============================================================================================
// LayerNorm
// =============================================================================================
template <typename T, int VEC>
__global__ void __launch_bounds__(128) ln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, T* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, long rows,
                                                     int d, float eps) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + row * d;
    float s = 0.f;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float v[VEC];
        VLoad<T, VEC>::ld(xr + c, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) s += v[i];
    }
    const float mu = warp_sum(s) / (float)d;
    float q = 0.f;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float v[VEC];
        VLoad<T, VEC>::ld(xr + c, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) { float t = v[i] - mu; q += t * t; }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)d + eps);
    T* yr = y + row * d;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float v[VEC], o[VEC];
        VLoad<T, VEC>::ld(xr + c, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = (v[i] - mu) * rs * gamma[c + i] + beta[c + i];
        VLoad<T, VEC>::st(yr + c, o);
    }
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(128) ln_bwd_dx_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const T* __restrict__ dres,
                                                        T* __restrict__ dx, long rows, int d) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + row * d;
    const T* gr = dy + row * d;
    const float mu = mean[row], rs = rstd[row];
    float a = 0.f, b = 0.f;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float v[VEC], g[VEC];
        VLoad<T, VEC>::ld(xr + c, v);
        VLoad<T, VEC>::ld(gr + c, g);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float gg = g[i] * gamma[c + i];
            a += gg;
            b += gg * (v[i] - mu) * rs;
        }
    }
    a = warp_sum(a) / (float)d;
    b = warp_sum(b) / (float)d;
    T* dr = dx + row * d;
    const T* rr = dres ? dres + row * d : nullptr;
    for (int c = lane * VEC; c < d; c += 32 * VEC) {
        float v[VEC], g[VEC], o[VEC];
        VLoad<T, VEC>::ld(xr + c, v);
        VLoad<T, VEC>::ld(gr + c, g);
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = rs * (g[i] * gamma[c + i] - a - (v[i] - mu) * rs * b);
        if (rr) {
            float e[VEC];
            VLoad<T, VEC>::ld(rr + c, e);
#pragma unroll
            for (int i = 0; i < VEC; ++i) o[i] += e[i];
        }
        VLoad<T, VEC>::st(dr + c, o);
    }
}


// ---------------------------------------------------------------------------------------------
// Row-in-registers LayerNorm (d % 8 == 0, d <= 2048): x (and dy) are read from HBM once with 16-byte loads.
// Rows stay PACKED in registers (4 registers per 8 bf16) and are decoded on each use: the register footprint, not the
// arithmetic, decides how many rows an SM keeps in flight.  dgamma / dbeta come from the 8-wide column reduction below
// (a register-accumulating fused variant was measured slower: 255 registers -> 8 warps per SM).
// ---------------------------------------------------------------------------------------------
// raw 16-byte row chunks: 8 bf16 in 4 registers (decoded on every use) or 8 floats in 8 registers
template <typename T, int NV>
__global__ void __launch_bounds__(256) ln_fwd_reg_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean,
                                                         float* __restrict__ rstd, long rows, int d, float eps) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + row * d;
    Raw8<T> raw[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < d) {
            float v[8];
            raw[j].load(xr + c);
            raw[j].get(v);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[k];
        }
    }
    const float mu = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < d) {
            float v[8];
            raw[j].get(v);
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float t = v[k] - mu; q += t * t; }
        }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)d + eps);
    T* yr = y + row * d;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < d) {
            float v[8], gm[8], bt[8], o[8];
            raw[j].get(v);
            Vec8<float>::load(gamma + c, gm);
            Vec8<float>::load(beta + c, bt);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = (v[k] - mu) * rs * gm[k] + bt[k];
            Vec8<T>::store(yr + c, o);
        }
    }
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

template <typename T, int NV>
__global__ void __launch_bounds__(256) ln_bwd_reg_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                         const float* __restrict__ gamma, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, const T* __restrict__ dres,
                                                         T* __restrict__ dx, long rows, int d, T* __restrict__ dxd, Dropout drop) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    if (dxd) dropout_resolve(drop);
    const T* xr = x + row * d;
    const T* gr = dy + row * d;
    const float mu = mean[row], rs = rstd[row];
    Raw8<T> rx[NV], rg[NV];
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < d) { rx[j].load(xr + c); rg[j].load(gr + c); }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < d) {
            float xv[8], gv[8], gm[8];
            rx[j].get(xv);
            rg[j].get(gv);
            Vec8<float>::load(gamma + c, gm);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float gg = gv[k] * gm[k];
                a += gg;
                b += gg * (xv[k] - mu) * rs;
            }
        }
    }
    a = warp_sum(a) / (float)d;
    b = warp_sum(b) / (float)d;
    T* dr = dx + row * d;
    const T* rr = dres ? dres + row * d : nullptr;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < d) {
            float xv[8], gv[8], gm[8], o[8];
            rx[j].get(xv);
            rg[j].get(gv);
            Vec8<float>::load(gamma + c, gm);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = rs * (gv[k] * gm[k] - a - (xv[k] - mu) * rs * b);
            if (rr) {
                float e[8];
                Vec8<T>::load(rr + c, e);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] += e[k];
            }
            Vec8<T>::store(dr + c, o);
            if (dxd) {                      // second output: the same gradient behind the dropout of the branch that fed the sum
                float mk[8];
                dropout_factors<8>(drop, (uint64_t)(row * d + c), mk);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] *= mk[k];
                Vec8<T>::store(dxd + row * d + c, o);
            }
        }
    }
}

// Same dX, plus the parameter gradients in the same pass: every warp walks rows blockIdx.x * 8 + warp, + gridDim.x * 8, ... and keeps
// its columns' sums of dy * xhat / dy in registers; the CTA folds its 8 warps through shared-memory atomics and issues ONE global
// float atomicAdd per (column, CTA).  Replaces ln_bwd_reg + a colreduce launch that re-read dy and x (rows > 512 only: below that the
// two-kernel route's single row chunk keeps the sums bit-reproducible, see launch_colreduce).
template <typename T, int NV>
__global__ void __launch_bounds__(256, NV <= 2 ? 3 : 1) ln_bwd_grad_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                          const float* __restrict__ gamma, const float* __restrict__ mean,
                                                          const float* __restrict__ rstd, const T* __restrict__ dres,
                                                          T* __restrict__ dx, long rows, int d, T* __restrict__ dxd, Dropout drop,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ float sacc[2][NV * 256];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * NV * 256; i += 256) (&sacc[0][0])[i] = 0.f;
    __syncthreads();
    if (dxd) dropout_resolve(drop);
    float ag[NV][8], ab[NV][8];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) { ag[j][k] = 0.f; ab[j][k] = 0.f; }
    for (long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long)gridDim.x * 8) {
        const T* xr = x + row * d;
        const T* gr = dy + row * d;
        const float mu = mean[row], rs = rstd[row];
        Raw8<T> rx[NV], rg[NV];
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 8;
            if (c < d) { rx[j].load(xr + c); rg[j].load(gr + c); }
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 8;
            if (c < d) {
                float xv[8], gv[8], gm[8];
                rx[j].get(xv);
                rg[j].get(gv);
                Vec8<float>::load(gamma + c, gm);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float xh = (xv[k] - mu) * rs, gg = gv[k] * gm[k];
                    a += gg;
                    b += gg * xh;
                    ag[j][k] += gv[k] * xh;
                    ab[j][k] += gv[k];
                }
            }
        }
        a = warp_sum(a) / (float)d;
        b = warp_sum(b) / (float)d;
        T* dr = dx + row * d;
        const T* rr = dres ? dres + row * d : nullptr;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int c = (j * 32 + lane) * 8;
            if (c < d) {
                float xv[8], gv[8], gm[8], o[8];
                rx[j].get(xv);
                rg[j].get(gv);
                Vec8<float>::load(gamma + c, gm);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = rs * (gv[k] * gm[k] - a - (xv[k] - mu) * rs * b);
                if (rr) {
                    float e[8];
                    Vec8<T>::load(rr + c, e);
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] += e[k];
                }
                Vec8<T>::store(dr + c, o);
                if (dxd) {
                    float mk[8];
                    dropout_factors<8>(drop, (uint64_t)(row * d + c), mk);
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] *= mk[k];
                    Vec8<T>::store(dxd + row * d + c, o);
                }
            }
        }
    }
    // shared layout [k][column group] so that a warp's 32 lanes hit 32 different banks
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int g = j * 32 + lane;
        if (g * 8 < d) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                atomicAdd(&sacc[0][k * (NV * 32) + g], ag[j][k]);
                atomicAdd(&sacc[1][k * (NV * 32) + g], ab[j][k]);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += 256) {
        const int i = (c & 7) * (NV * 32) + (c >> 3);
        atomicAdd(dgamma + c, sacc[0][i]);
        atomicAdd(dbeta + c, sacc[1][i]);
    }
}

// =============================================================================================
// Generic column reduction: out_k[c] += sum_r f_k(r, c).  blockDim = (32, 8); each thread owns
// VEC consecutive columns; grid = (col tiles, row chunks); cross-warp reduce in shared memory and
// one float atomicAdd per (k, column, row chunk).
// =============================================================================================
template <int NOUT, int VEC, typename F>
__global__ void __launch_bounds__(256) colreduce_kernel(F f, long rows, int cols) {
    __shared__ float red[8][NOUT][32 * VEC];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c0 = (blockIdx.x * 32 + tx) * VEC;
    float acc[NOUT][VEC];
#pragma unroll
    for (int k = 0; k < NOUT; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[k][v] = 0.f;
    const long per = (rows + gridDim.y - 1) / gridDim.y;
    const long r0 = (long)blockIdx.y * per;
    const long r1 = (r0 + per < rows) ? r0 + per : rows;
    if (c0 < cols) {
#pragma unroll(F::kUnroll)
        for (long r = r0 + ty; r < r1; r += 8) f(r, c0, acc);
    }
#pragma unroll
    for (int k = 0; k < NOUT; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) red[ty][k][tx * VEC + v] = acc[k][v];
    __syncthreads();
    // 256 threads sum NOUT * 32 * VEC entries over the 8 row-warps
    for (int e = ty * 32 + tx; e < NOUT * 32 * VEC; e += 256) {
        int k = e / (32 * VEC), cc = e % (32 * VEC);
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][k][cc];
        int c = blockIdx.x * 32 * VEC + cc;
        if (c < cols) atomicAdd(f.out(k, c), s);
    }
}

template <int NOUT, int VEC, typename F>
static int launch_colreduce(F f, long rows, int cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return S2S_OK;
    unsigned gx = (unsigned)ceil_div_l(cols, 32 * VEC);
    long want = (long)num_sms() * 4 / gx;
    if (want < 1) want = 1;
    long maxy = ceil_div_l(rows, 64);
    if (want > maxy) want = maxy;
    if (rows <= 512) want = 1;      // small problems: one row chunk per column tile, so no floating-point atomics meet and the
                                    // sums (BatchNorm statistics feed the forward pass) are bit-reproducible run to run
    if (want < 1) want = 1;
    colreduce_kernel<NOUT, VEC, F><<<dim3(gx, (unsigned)want), dim3(32, 8), 0, st>>>(f, rows, cols);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

template <typename T, int VEC> struct ColsumF {
    static constexpr int kUnroll = 4;   // rows in flight per thread in colreduce_kernel
    const T* x; long ld; float* o;
    __device__ __forceinline__ void operator()(long r, int c0, float (&acc)[1][VEC]) const {
        float v[VEC];
        VLoad<T, VEC>::ld(x + r * ld + c0, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[0][i] += v[i];
    }
    __device__ __forceinline__ float* out(int, int c) const { return o + c; }
};

template <typename T, int VEC> struct LNGradF {
    static constexpr int kUnroll = 4;   // rows in flight per thread in colreduce_kernel
    const T* dy; const T* x; const float* mean; const float* rstd; int d; float* dgamma; float* dbeta;
    __device__ __forceinline__ void operator()(long r, int c0, float (&acc)[2][VEC]) const {
        float v[VEC], g[VEC];
        VLoad<T, VEC>::ld(x + r * d + c0, v);
        VLoad<T, VEC>::ld(dy + r * d + c0, g);
        const float mu = mean[r], rs = rstd[r];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { acc[0][i] += g[i] * (v[i] - mu) * rs; acc[1][i] += g[i]; }
    }
    __device__ __forceinline__ float* out(int k, int c) const { return (k == 0 ? dgamma : dbeta) + c; }
};

// =============================================================================================
// elementwise: relu backward, dropout backward
// =============================================================================================
template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, long n,
                                float scale) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        dx[i] = from_f<T>(to_f<T>(y[i]) > 0.f ? to_f<T>(dy[i]) * scale : 0.f);
}
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd8_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx,
                                                        long n8, float scale) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        float g[8], a[8];
        Vec8<T>::load(dy + 8 * i, g);
        Vec8<T>::load(y + 8 * i, a);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = a[k] > 0.f ? g[k] * scale : 0.f;
        Vec8<T>::store(dx + 8 * i, g);
    }
}
template <typename T>
__global__ void dropout_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, long n, Dropout drop) {
    dropout_resolve(drop);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        dx[i] = from_f<T>(to_f<T>(dy[i]) * dropout_factor(drop, (uint64_t)i));
}
template <typename T>
__global__ void __launch_bounds__(256) dropout_bwd8_kernel(const T* __restrict__ dy, T* __restrict__ dx, long n8, Dropout drop) {
    dropout_resolve(drop);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        float g[8];
        Vec8<T>::load(dy + 8 * i, g);
        float mk[8];
        dropout_factors<8>(drop, (uint64_t)i * 8, mk);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] *= mk[k];
        Vec8<T>::store(dx + 8 * i, g);
    }
}

// =============================================================================================
// BatchNorm over (B, L) frames of a haloed (B, Lp = L + 2*halo, C) buffer
// =============================================================================================
struct BNGeom {
    int L, Lp, halo, C;
    __device__ __forceinline__ long phys(long r) const {  // frame index -> physical row
        if (halo == 0) return r;
        if (r < 0x7fffffffL) {
            const unsigned ru = (unsigned)r, b = ru / (unsigned)L;
            return (long)b * Lp + halo + (int)(ru - b * (unsigned)L);
        }
        long b = r / L;
        int l = (int)(r - b * L);
        return b * Lp + halo + l;
    }
};

template <typename T, int VEC> struct BNStatsF {
    static constexpr int kUnroll = 4;   // rows in flight per thread in colreduce_kernel
    const T* x; BNGeom g; float* sums;
    __device__ __forceinline__ void operator()(long r, int c0, float (&acc)[2][VEC]) const {
        float v[VEC];
        VLoad<T, VEC>::ld(x + g.phys(r) * g.C + c0, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) { acc[0][i] += v[i]; acc[1][i] += v[i] * v[i]; }
    }
    __device__ __forceinline__ float* out(int k, int c) const { return sums + k * g.C + c; }
};

__global__ void bn_finalize_kernel(const float* __restrict__ sums, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* running_mean, float* running_var, long count,
                                   int C, float eps, float momentum) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float n = (float)count;
    float mu = sums[c] / n;
    float var = sums[C + c] / n - mu * mu;
    var = fmaxf(var, 0.f);
    mean[c] = mu;
    invstd[c] = rsqrtf(var + eps);
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
    if (running_var) {
        float unb = count > 1 ? var * n / (n - 1.f) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv,
                                     float* __restrict__ mean, float* __restrict__ invstd, int C, float eps) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = rm[c];
    invstd[c] = rsqrtf(rv[c] + eps);
}

template <typename T, int VEC>
__global__ void bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ mean,
                                const float* __restrict__ invstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, T* __restrict__ y, int B, BNGeom g, int use_tanh,
                                Dropout drop) {
    dropout_resolve(drop);
    const int cv = g.C / VEC;
    const long total = (long)B * g.Lp * cv;
    const bool small = total < 0x7fffffffL;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c, lp;
        long prow, b;
        if (small) {                                    // 32-bit divisions: the 64-bit ones cost more than the memory traffic
            const unsigned iu = (unsigned)i, pr = iu / (unsigned)cv;
            c = (int)(iu - pr * (unsigned)cv) * VEC;
            const unsigned bb = pr / (unsigned)g.Lp;
            lp = (int)(pr - bb * (unsigned)g.Lp);
            prow = pr; b = bb;
        } else {
            c = (int)(i % cv) * VEC;
            prow = i / cv;
            lp = (int)(prow % g.Lp);
            b = prow / g.Lp;
        }
        float o[VEC];
        int l = lp - g.halo;
        if (l < 0 || l >= g.L) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = 0.f;
        } else {
            float v[VEC];
            VLoad<T, VEC>::ld(x + prow * g.C + c, v);
            uint64_t idx = (uint64_t)((b * g.L + l) * (long)g.C + c);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float t = (v[k] - mean[c + k]) * invstd[c + k] * gamma[c + k] + beta[c + k];
                if (use_tanh == 1) t = tanhf(t);
                else if (use_tanh == 2) t = t / (1.f + __expf(-t));      // Swish (conformer/convolution.py:75)
                o[k] = t * dropout_factor(drop, idx + k);
            }
        }
        VLoad<T, VEC>::st(y + prow * g.C + c, o);
    }
}

// dz = dy * dropmask * act'(y_pre_dropout); y holds the post-activation, post-dropout output.
// With dropout the pre-dropout activation is recomputed from x so tanh' stays exact.
template <typename T, int VEC>
__device__ __forceinline__ void bn_dz(const T* dy, const T* y, const T* x, const float* mean, const float* invstd,
                                      const float* gamma, const float* beta, const BNGeom& g, long r, int c0,
                                      int use_tanh, const Dropout& drop, float (&dz)[VEC], float (&xhat)[VEC]) {
    long prow = g.phys(r);
    float gy[VEC], xv[VEC];
    VLoad<T, VEC>::ld(dy + prow * g.C + c0, gy);
    VLoad<T, VEC>::ld(x + prow * g.C + c0, xv);
    uint64_t idx = (uint64_t)(r * (long)g.C + c0);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        xhat[k] = (xv[k] - mean[c0 + k]) * invstd[c0 + k];
        float t = gy[k] * dropout_factor(drop, idx + k);
        if (use_tanh == 1) {
            float a = tanhf(xhat[k] * gamma[c0 + k] + beta[c0 + k]);
            t *= (1.f - a * a);
        } else if (use_tanh == 2) {
            float z = xhat[k] * gamma[c0 + k] + beta[c0 + k];
            float sg = 1.f / (1.f + __expf(-z));
            t *= sg * (1.f + z * (1.f - sg));
        }
        dz[k] = t;
    }
}

template <typename T, int VEC> struct BNBwdF {
    static constexpr int kUnroll = 1;   // rows in flight per thread in colreduce_kernel
    const T* dy; const T* y; const T* x; const float* mean; const float* invstd; const float* gamma;
    const float* beta; BNGeom g; int use_tanh; Dropout drop; float* sums;
    __device__ __forceinline__ void operator()(long r, int c0, float (&acc)[2][VEC]) const {
        float dz[VEC], xh[VEC];
        Dropout d = drop;
        dropout_resolve(d);
        bn_dz<T, VEC>(dy, y, x, mean, invstd, gamma, beta, g, r, c0, use_tanh, d, dz, xh);
#pragma unroll
        for (int k = 0; k < VEC; ++k) { acc[0][k] += dz[k]; acc[1][k] += dz[k] * xh[k]; }
    }
    __device__ __forceinline__ float* out(int k, int c) const { return sums + k * g.C + c; }
};

template <typename T, int VEC>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ sums, T* __restrict__ dx, int B, BNGeom g,
                                    int use_tanh, Dropout drop) {
    dropout_resolve(drop);
    const int cv = g.C / VEC;
    const long total = (long)B * g.Lp * cv;
    const float inv_n = 1.f / (float)((long)B * g.L);
    const bool small = total < 0x7fffffffL;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c, lp;
        long prow, b;
        if (small) {                                    // 32-bit divisions: the 64-bit ones cost more than the memory traffic
            const unsigned iu = (unsigned)i, pr = iu / (unsigned)cv;
            c = (int)(iu - pr * (unsigned)cv) * VEC;
            const unsigned bb = pr / (unsigned)g.Lp;
            lp = (int)(pr - bb * (unsigned)g.Lp);
            prow = pr; b = bb;
        } else {
            c = (int)(i % cv) * VEC;
            prow = i / cv;
            lp = (int)(prow % g.Lp);
            b = prow / g.Lp;
        }
        int l = lp - g.halo;
        float o[VEC];
        if (l < 0 || l >= g.L) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) o[k] = 0.f;
        } else {
            float dz[VEC], xh[VEC];
            bn_dz<T, VEC>(dy, y, x, mean, invstd, gamma, beta, g, b * g.L + l, c, use_tanh, drop, dz, xh);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float t = dz[k];
                if (sums) t -= sums[c + k] * inv_n + xh[k] * sums[g.C + c + k] * inv_n;
                o[k] = t * gamma[c + k] * invstd[c + k];
            }
        }
        VLoad<T, VEC>::st(dx + prow * g.C + c, o);
    }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm apply / backward-apply without halos (the Conformer convolution module): flat 16-byte grid-stride kernels with the
// per-channel constants staged once per CTA in shared memory and the channel index advanced incrementally (no division, no
// per-element parameter loads from global memory), few registers so that many CTAs keep enough bytes in flight.
// ---------------------------------------------------------------------------------------------
template <int ACT> __device__ __forceinline__ float bn_act(float t) {
    if (ACT == 1) return tanhf(t);
    if (ACT == 2) return __fdividef(t, 1.f + __expf(-t));          // Swish (conformer/convolution.py:75)
    return t;
}
template <int ACT> __device__ __forceinline__ float bn_act_grad(float z) {
    if (ACT == 1) { const float a = tanhf(z); return 1.f - a * a; }
    if (ACT == 2) { const float sg = __fdividef(1.f, 1.f + __expf(-z)); return sg * (1.f + z * (1.f - sg)); }
    return 1.f;
}

// position of a thread's current 8-channel vector inside a (B, Lp = L + 2 * halo, C) buffer, advanced by the grid stride with
// carries instead of divisions
template <bool HALO> struct BNCursor;
template <> struct BNCursor<true> {
    int ci, lp, b;
    int dci, dlp, db;            // the grid stride split into (channel vectors, rows inside an utterance, utterances)
    __device__ __forceinline__ void init(long i, long step, int cv, int Lp) {
        long row = i / cv; ci = (int)(i - row * cv); b = (int)(row / Lp); lp = (int)(row - (long)b * Lp);
        long srow = step / cv; dci = (int)(step - srow * cv); db = (int)(srow / Lp); dlp = (int)(srow - (long)db * Lp);
    }
    __device__ __forceinline__ void next(int cv, int Lp) {
        ci += dci; lp += dlp; b += db;
        if (ci >= cv) { ci -= cv; ++lp; }
        if (lp >= Lp) { lp -= Lp; ++b; }
    }
    __device__ __forceinline__ bool frame(const BNGeom& g, long, uint64_t& idx) const {   // element index of the frame (dropout counter)
        const int l = lp - g.halo;
        idx = (uint64_t)(((long)b * g.L + l) * (long)g.C + ci * 8);
        return l >= 0 && l < g.L;
    }
};
template <> struct BNCursor<false> {                       // no halos: physical rows are frames, only the channel vector is tracked
    int ci, dci;
    __device__ __forceinline__ void init(long i, long step, int cv, int) { ci = (int)(i % cv); dci = (int)(step % cv); }
    __device__ __forceinline__ void next(int cv, int) { ci += dci; if (ci >= cv) ci -= cv; }
    __device__ __forceinline__ bool frame(const BNGeom&, long i, uint64_t& idx) const { idx = (uint64_t)i * 8; return true; }
};

template <typename T, int ACT, bool HALO>
__global__ void __launch_bounds__(256) bn_apply_flat_kernel(const T* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, T* __restrict__ y, long nvec, BNGeom g,
                                                            Dropout drop) {
    extern __shared__ float bnp[];                     // a[C] = invstd * gamma, sh[C] = beta - mean * a
    dropout_resolve(drop);
    const int C = g.C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float a = invstd[c] * gamma[c];
        bnp[c] = a;
        bnp[C + c] = beta[c] - mean[c] * a;
    }
    __syncthreads();
    const int cv = C >> 3;
    const long step = (long)gridDim.x * blockDim.x;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    BNCursor<HALO> cur;
    cur.init(i, step, cv, g.Lp);
    for (; i < nvec; i += step, cur.next(cv, g.Lp)) {
        float o[8];
        uint64_t idx;
        if (cur.frame(g, i, idx)) {
            float v[8], mk[8];
            Vec8<T>::load(x + i * 8, v);
            const float* pa = bnp + cur.ci * 8;
            const float4 a0 = *reinterpret_cast<const float4*>(pa), a1 = *reinterpret_cast<const float4*>(pa + 4);
            const float4 s0 = *reinterpret_cast<const float4*>(pa + C), s1 = *reinterpret_cast<const float4*>(pa + C + 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            dropout_factors<8>(drop, idx, mk);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = bn_act<ACT>(fmaf(v[k], a[k], sh[k])) * mk[k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = 0.f;         // halo rows of the output stay zero
        }
        Vec8<T>::store(y + i * 8, o);
    }
}

// dz = dy * dropmask * act'(xhat * gamma + beta) and xhat for one 8-channel vector; pc -> {mean, invstd, gamma, beta, ...}[C]
template <int ACT>
__device__ __forceinline__ void bn_dz_flat(const float (&gy)[8], const float (&xv)[8], const float* pc, int C, const Dropout& drop,
                                           uint64_t idx, float (&pr)[4][8], float (&dz)[8], float (&xh)[8]) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float4 lo = *reinterpret_cast<const float4*>(pc + a * C), hi = *reinterpret_cast<const float4*>(pc + a * C + 4);
        pr[a][0] = lo.x; pr[a][1] = lo.y; pr[a][2] = lo.z; pr[a][3] = lo.w; pr[a][4] = hi.x; pr[a][5] = hi.y; pr[a][6] = hi.z; pr[a][7] = hi.w;
    }
    float mk[8];
    dropout_factors<8>(drop, idx, mk);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        xh[k] = (xv[k] - pr[0][k]) * pr[1][k];
        dz[k] = gy[k] * mk[k] * bn_act_grad<ACT>(fmaf(xh[k], pr[2][k], pr[3][k]));
    }
}

template <typename T, int ACT, bool HALO>
__global__ void __launch_bounds__(256) bn_bwd_apply_flat_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const float* __restrict__ sums, T* __restrict__ dx, long nvec, BNGeom g,
                                                                float inv_n, Dropout drop) {
    extern __shared__ float bnp[];                     // mean | invstd | gamma | beta | sums0 / n | sums1 / n
    dropout_resolve(drop);
    const int C = g.C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        bnp[c] = mean[c]; bnp[C + c] = invstd[c]; bnp[2 * C + c] = gamma[c]; bnp[3 * C + c] = beta[c];
        bnp[4 * C + c] = sums ? sums[c] * inv_n : 0.f;
        bnp[5 * C + c] = sums ? sums[C + c] * inv_n : 0.f;
    }
    __syncthreads();
    const int cv = C >> 3;
    const long step = (long)gridDim.x * blockDim.x;
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    BNCursor<HALO> cur;
    cur.init(i, step, cv, g.Lp);
    for (; i < nvec; i += step, cur.next(cv, g.Lp)) {
        float o[8];
        uint64_t idx;
        if (cur.frame(g, i, idx)) {
            float gy[8], xv[8], pr[4][8], dz[8], xh[8];
            Vec8<T>::load(dy + i * 8, gy);
            Vec8<T>::load(x + i * 8, xv);
            const float* pc = bnp + cur.ci * 8;
            bn_dz_flat<ACT>(gy, xv, pc, C, drop, idx, pr, dz, xh);
            const float4 p0 = *reinterpret_cast<const float4*>(pc + 4 * C), p1 = *reinterpret_cast<const float4*>(pc + 4 * C + 4);
            const float4 q0 = *reinterpret_cast<const float4*>(pc + 5 * C), q1 = *reinterpret_cast<const float4*>(pc + 5 * C + 4);
            const float s0[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w}, s1[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = (dz[k] - s0[k] - xh[k] * s1[k]) * pr[2][k] * pr[1][k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = 0.f;
        }
        Vec8<T>::store(dx + i * 8, o);
    }
}

static inline unsigned bn_flat_grid(long nvec) {          // every CTA stages the per-channel constants: keep the grid near-persistent
    long b = ceil_div_l(nvec, 256 * 4);
    const long cap = (long)num_sms() * 6;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

#define S2S_BN_ACT_DISPATCH2(act, ACT, ...)                \
    do {                                                    \
        if ((act) == 1) { constexpr int ACT = 1; __VA_ARGS__; }      \
        else if ((act) == 2) { constexpr int ACT = 2; __VA_ARGS__; } \
        else { constexpr int ACT = 0; __VA_ARGS__; }                 \
    } while (0)
#define S2S_BN_ACT_DISPATCH(act, ACT, ...)                 \
    do {                                                    \
        if (halo > 0) { constexpr bool HALO = true; S2S_BN_ACT_DISPATCH2(act, ACT, __VA_ARGS__); }  \
        else { constexpr bool HALO = false; S2S_BN_ACT_DISPATCH2(act, ACT, __VA_ARGS__); }          \
    } while (0)

__global__ void bn_param_grad_kernel(const float* __restrict__ sums, float* dgamma, float* dbeta, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    if (dbeta) dbeta[c] += sums[c];
    if (dgamma) dgamma[c] += sums[C + c];
}


template <typename T, int VEC>
__global__ void pad_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int L, int halo, int C, int to_padded) {
    const int cv = C / VEC, Lp = L + 2 * halo;
    const long total = (long)B * (to_padded ? Lp : L) * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % cv) * VEC;
        long row = i / cv;
        float v[VEC];
        if (to_padded) {
            int lp = (int)(row % Lp);
            long b = row / Lp;
            int l = lp - halo;
            if (l < 0 || l >= L) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) v[k] = 0.f;
            } else {
                VLoad<T, VEC>::ld(x + (b * L + l) * C + c, v);
            }
            VLoad<T, VEC>::st(y + row * C + c, v);
        } else {
            int l = (int)(row % L);
            long b = row / L;
            VLoad<T, VEC>::ld(x + (b * Lp + halo + l) * C + c, v);
            VLoad<T, VEC>::st(y + row * C + c, v);
        }
    }
}


// =============================================================================================
// Skinny linear layer (N <= 4 output features; the stop-token head prob_out, models/vtn.py:182,251).
// A tensor-core tile would be > 95 % padding here and the weight-gradient GEMM has a single output
// tile, so these are bandwidth-shaped kernels: every row of x is read once per pass.
// =============================================================================================
template <typename T, int VEC>
__global__ void __launch_bounds__(128) skinny_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w,
                                                         const float* __restrict__ bias, T* __restrict__ y, long rows,
                                                         int K, int N) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* xr = x + row * K;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = lane * VEC; c < K; c += 32 * VEC) {
        float v[VEC];
        VLoad<T, VEC>::ld(xr + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < N) {
                float wv[VEC];
                VLoad<T, VEC>::ld(w + (long)j * K + c, wv);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[j] = fmaf(v[i], wv[i], acc[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
    if (lane < N) {
        float o = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        y[row * N + lane] = from_f<T>(o + (bias ? bias[lane] : 0.f));
    }
}

template <typename T, int VEC> struct SkinnyDwF {
    static constexpr int kUnroll = 1;   // rows in flight per thread in colreduce_kernel
    const T* dy; const T* x; int K; int N; float* dw;
    __device__ __forceinline__ void operator()(long r, int c0, float (&acc)[4][VEC]) const {
        float v[VEC];
        VLoad<T, VEC>::ld(x + r * K + c0, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < N) {
                const float g = to_f<T>(dy[r * N + j]);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[j][i] = fmaf(g, v[i], acc[j][i]);
            }
        }
    }
    __device__ __forceinline__ float* out(int k, int c) const { return dw + (long)(k < N ? k : 0) * K + c; }
};

template <typename T, int VEC>
__global__ void skinny_dx_kernel(const T* __restrict__ dy, const T* __restrict__ w, T* __restrict__ dx, long rows, int K,
                                 int N, int accumulate) {
    const int kv = K / VEC;
    const long total = rows * kv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / kv;
        const int c = (int)(i - r * kv) * VEC;
        float o[VEC];
        if (accumulate) VLoad<T, VEC>::ld(dx + r * K + c, o);
        else {
#pragma unroll
            for (int q = 0; q < VEC; ++q) o[q] = 0.f;
        }
        for (int j = 0; j < N; ++j) {
            const float g = to_f<T>(dy[r * N + j]);
            float wv[VEC];
            VLoad<T, VEC>::ld(w + (long)j * K + c, wv);
#pragma unroll
            for (int q = 0; q < VEC; ++q) o[q] = fmaf(g, wv[q], o[q]);
        }
        VLoad<T, VEC>::st(dx + r * K + c, o);
    }
}

}  // namespace s2s

using namespace s2s;

#define S2S_VEC3_DISPATCH(ok8, ok4, VEC, ...)              \
    do {                                                   \
        if (ok8) { constexpr int VEC = 8; __VA_ARGS__; }   \
        else if (ok4) { constexpr int VEC = 4; __VA_ARGS__; } \
        else { constexpr int VEC = 1; __VA_ARGS__; }       \
    } while (0)

#define S2S_VEC_DISPATCH(ok, VEC, ...)                \
    do {                                              \
        if (ok) { constexpr int VEC = 4; __VA_ARGS__; } \
        else { constexpr int VEC = 1; __VA_ARGS__; }  \
    } while (0)

extern "C" int s2s_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                 float* rstd, int64_t rows, int d, float eps, int dtype, void* stream) {
    S2S_REQUIRE(x && gamma && beta && y && mean && rstd && d > 0, "layernorm_fwd: null pointer or bad d");
    if (rows <= 0) return S2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (d % 8 == 0 && d <= 2048 && aligned16(x, y, gamma, beta)) {
        const int nv = (d + 255) / 256;
        unsigned g8 = (unsigned)ceil_div_l(rows, 8);
#define S2S_LN_FWD(NVV) ln_fwd_reg_kernel<T, NVV><<<g8, 256, 0, st>>>((const T*)x, gamma, beta, (T*)y, mean, rstd, rows, d, eps)
        S2S_DISPATCH_DTYPE(dtype, T, {
            if (nv <= 1) S2S_LN_FWD(1); else if (nv <= 2) S2S_LN_FWD(2); else if (nv <= 4) S2S_LN_FWD(4);
            else if (nv <= 6) S2S_LN_FWD(6); else S2S_LN_FWD(8);
        });
#undef S2S_LN_FWD
        S2S_LAUNCH_OK();
        return S2S_OK;
    }
    bool ok = vec4_ok(d, d, x, y);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, (ln_fwd_kernel<T, VEC><<<(unsigned)ceil_div_l(rows, 4), 128, 0, st>>>(
        (const T*)x, gamma, beta, (T*)y, mean, rstd, rows, d, eps))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

// S2S_LN_FUSED=0 restores the two-kernel route (A/B measurements)
static int g_ln_fused = [] { const char* e = getenv("S2S_LN_FUSED"); return e ? atoi(e) : 1; }();

static int layernorm_bwd_impl(const void* dy, const void* x, const float* gamma, const float* mean,
                                 const float* rstd, const void* dres, void* dx, float* dgamma, float* dbeta,
                                 int64_t rows, int d, int dtype, void* stream, void* dx_drop, const s2s_dropout_t* dropd) {
    S2S_REQUIRE(dy && x && gamma && mean && rstd && d > 0, "layernorm_bwd: null pointer or bad d");
    if (rows <= 0) return S2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (d % 8 == 0 && d <= 2048 && aligned16(x, dy, dx, dres) && aligned16(gamma, dx_drop)) {
        const int nv = (d + 255) / 256;
        const unsigned gp = (unsigned)ceil_div_l(rows, 8);
#define S2S_LN_BWD(NVV) ln_bwd_reg_kernel<T, NVV><<<gp, 256, 0, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, \
                                                                     (T*)dx, rows, d, (T*)dx_drop, dd)
        const Dropout dd = make_dropout(dropd);
        if (dx && dgamma && dbeta && rows > 512 && nv <= 2 && g_ln_fused) {
            // one pass: dX and the gamma / beta gradients (3 resident CTAs per SM, every CTA loops over its share of the rows)
            const long groups = ceil_div_l(rows, 8);
            const long cap = (long)num_sms() * 3;
            const unsigned gf = (unsigned)(groups < cap ? groups : cap);
#define S2S_LN_BWDG(NVV) ln_bwd_grad_kernel<T, NVV><<<gf, 256, 0, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, \
                                                                      (T*)dx, rows, d, (T*)dx_drop, dd, dgamma, dbeta)
            S2S_DISPATCH_DTYPE(dtype, T, {
                if (nv <= 1) S2S_LN_BWDG(1); else S2S_LN_BWDG(2);
            });
#undef S2S_LN_BWDG
            S2S_LAUNCH_OK();
            return S2S_OK;
        }
        if (dx) {
            S2S_DISPATCH_DTYPE(dtype, T, {
                if (nv <= 1) S2S_LN_BWD(1); else if (nv <= 2) S2S_LN_BWD(2); else if (nv <= 4) S2S_LN_BWD(4);
                else if (nv <= 6) S2S_LN_BWD(6); else S2S_LN_BWD(8);
            });
            S2S_LAUNCH_OK();
        }
#undef S2S_LN_BWD
        if (dgamma && dbeta) {
            int rc = S2S_OK;
            S2S_DISPATCH_DTYPE(dtype, T, {
                LNGradF<T, 8> f{(const T*)dy, (const T*)x, mean, rstd, d, dgamma, dbeta};
                rc = launch_colreduce<2, 8>(f, rows, d, st);
            });
            return rc;
        }
        return S2S_OK;
    }
    bool ok = vec4_ok(d, d, x, dy, dx, dres);
    if (dx_drop && !dx) return set_error(S2S_ERR_INVALID, "layernorm_bwd_drop: dx_drop needs dx");
    if (dx) {
        S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, (ln_bwd_dx_kernel<T, VEC><<<(unsigned)ceil_div_l(rows, 4), 128, 0, st>>>(
            (const T*)dy, (const T*)x, gamma, mean, rstd, (const T*)dres, (T*)dx, rows, d))));
        S2S_LAUNCH_OK();
        if (dx_drop) {
            int rc = s2s_dropout_bwd(dx, dx_drop, rows, d, dropd, dtype, stream);
            if (rc != S2S_OK) return rc;
        }
    }
    if (dgamma && dbeta) {
        int rc = S2S_OK;
        S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, {
            LNGradF<T, VEC> f{(const T*)dy, (const T*)x, mean, rstd, d, dgamma, dbeta};
            rc = launch_colreduce<2, VEC>(f, rows, d, st);
        }));
        return rc;
    }
    return S2S_OK;
}

extern "C" int s2s_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                 const void* dres, void* dx, float* dgamma, float* dbeta, int64_t rows, int d, int dtype, void* stream) {
    return layernorm_bwd_impl(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, d, dtype, stream, nullptr, nullptr);
}

extern "C" int s2s_layernorm_bwd_drop(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                                      const void* dres, void* dx, void* dx_drop, const s2s_dropout_t* drop, float* dgamma, float* dbeta,
                                      int64_t rows, int d, int dtype, void* stream) {
    S2S_REQUIRE(dx && dx_drop && drop, "layernorm_bwd_drop: dx, dx_drop and drop are required");
    return layernorm_bwd_impl(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, d, dtype, stream, dx_drop, drop);
}

extern "C" int s2s_colsum(const void* x, int64_t rows, int cols, int64_t ld, float* out, int dtype, void* stream) {
    S2S_REQUIRE(x && out && cols > 0 && ld >= cols, "colsum: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = S2S_OK;
    if (cols % 8 == 0 && ld % 8 == 0 && aligned16(x)) {
        S2S_DISPATCH_DTYPE(dtype, T, {
            ColsumF<T, 8> f{(const T*)x, (long)ld, out};
            rc = launch_colreduce<1, 8>(f, rows, cols, st);
        });
        return rc;
    }
    bool ok = vec4_ok(cols, ld, x);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, {
        ColsumF<T, VEC> f{(const T*)x, (long)ld, out};
        rc = launch_colreduce<1, VEC>(f, rows, cols, st);
    }));
    return rc;
}

// ---------------------------------------------------------------------------------------------
// several column sums in one launch (the bias gradients of one Transformer layer's backward)
// ---------------------------------------------------------------------------------------------
namespace s2s {
constexpr int COLSUM_MAX = 8;
struct ColsumSet {
    s2s_colsum_t it[COLSUM_MAX];
};
template <typename T>
__global__ void __launch_bounds__(256) colsum_multi_kernel(const __grid_constant__ ColsumSet set) {
    __shared__ float red[8][256];
    const s2s_colsum_t& d = set.it[blockIdx.z];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c0 = (blockIdx.x * 32 + tx) * 8;
    float acc[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[v] = 0.f;
    const long per = (d.rows + gridDim.y - 1) / gridDim.y;
    const long r0 = (long)blockIdx.y * per;
    const long r1 = (r0 + per < d.rows) ? r0 + per : d.rows;
    if (c0 < d.cols) {
        const T* x = reinterpret_cast<const T*>(d.x);
#pragma unroll 4
        for (long r = r0 + ty; r < r1; r += 8) {
            float v[8];
            Vec8<T>::load(x + r * d.ld + c0, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += v[i];
        }
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) red[ty][tx * 8 + v] = acc[v];
    __syncthreads();
    const int e = ty * 32 + tx;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][e];
    const int c = blockIdx.x * 256 + e;
    if (c < d.cols && r0 < r1) atomicAdd(d.out + c, sum);
}
}  // namespace s2s

extern "C" int s2s_colsum_multi(const s2s_colsum_t* items, int n, int dtype, void* stream) {
    S2S_REQUIRE(items && n >= 1, "colsum_multi: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int i = 0;
    while (i < n) {
        // pack runs of vectorisable items into one launch; anything else goes through s2s_colsum
        ColsumSet set;
        int m = 0, max_cols = 0;
        long max_rows = 0;
        while (i < n && m < COLSUM_MAX) {
            const s2s_colsum_t& d = items[i];
            S2S_REQUIRE(d.x && d.out && d.cols > 0 && d.ld >= d.cols, "colsum_multi: bad item %d", i);
            if (!(d.cols % 8 == 0 && d.ld % 8 == 0 && aligned16(d.x))) break;
            set.it[m++] = d;
            max_cols = d.cols > max_cols ? d.cols : max_cols;
            max_rows = d.rows > max_rows ? d.rows : max_rows;
            ++i;
        }
        if (m > 0 && max_rows > 0) {
            const unsigned gx = (unsigned)ceil_div_l(max_cols, 256);
            long want = (long)num_sms() * 4 / ((long)gx * m);
            const long maxy = ceil_div_l(max_rows, 64);
            if (want > maxy) want = maxy;
            if (want < 1) want = 1;
            S2S_DISPATCH_DTYPE(dtype, T, (colsum_multi_kernel<T><<<dim3(gx, (unsigned)want, (unsigned)m), dim3(32, 8), 0, st>>>(set)));
            S2S_LAUNCH_OK();
        }
        if (i < n && m < COLSUM_MAX) {          // the item that broke the run
            const s2s_colsum_t& d = items[i];
            const int rc = s2s_colsum(d.x, d.rows, d.cols, d.ld, d.out, dtype, stream);
            if (rc != S2S_OK) return rc;
            ++i;
        }
    }
    return S2S_OK;
}

extern "C" int s2s_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, float scale, int dtype, void* stream) {
    S2S_REQUIRE(dy && y && dx, "relu_bwd: null pointer");
    if (n <= 0) return S2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (n % 8 == 0 && aligned16(dy, y, dx)) {
        S2S_DISPATCH_DTYPE(dtype, T, (relu_bwd8_kernel<T><<<ew_grid(n / 8, 256), 256, 0, st>>>((const T*)dy, (const T*)y, (T*)dx, n / 8, scale)));
    } else {
        S2S_DISPATCH_DTYPE(dtype, T, (relu_bwd_kernel<T><<<ew_grid(n, 256 * 4), 256, 0, st>>>((const T*)dy, (const T*)y, (T*)dx, n, scale)));
    }
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_dropout_bwd(const void* dy, void* dx, int64_t rows, int cols, const s2s_dropout_t* drop, int dtype,
                               void* stream) {
    S2S_REQUIRE(dy && dx, "dropout_bwd: null pointer");
    long n = (long)rows * cols;
    if (n <= 0) return S2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Dropout d = make_dropout(drop);
    if (n % 8 == 0 && aligned16(dy, dx)) {
        S2S_DISPATCH_DTYPE(dtype, T, (dropout_bwd8_kernel<T><<<ew_grid(n / 8, 256), 256, 0, st>>>((const T*)dy, (T*)dx, n / 8, d)));
    } else {
        S2S_DISPATCH_DTYPE(dtype, T, (dropout_bwd_kernel<T><<<ew_grid(n, 256 * 4), 256, 0, st>>>((const T*)dy, (T*)dx, n, d)));
    }
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_bn_stats(const void* x, float* sums, int B, int L, int halo, int C, int dtype, void* stream) {
    S2S_REQUIRE(x && sums && B > 0 && L > 0 && C > 0 && halo >= 0, "bn_stats: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    BNGeom g{L, L + 2 * halo, halo, C};
    bool ok = vec4_ok(C, C, x), ok8 = (C % 8 == 0) && aligned16(x);
    int rc = S2S_OK;
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC3_DISPATCH(ok8, ok, VEC, {
        BNStatsF<T, VEC> f{(const T*)x, g, sums};
        rc = launch_colreduce<2, VEC>(f, (long)B * L, C, st);
    }));
    return rc;
}

extern "C" int s2s_bn_finalize(const float* sums, float* mean, float* invstd, float* running_mean,
                               float* running_var, int64_t count, int C, float eps, float momentum, void* stream) {
    S2S_REQUIRE(sums && mean && invstd && count > 0 && C > 0, "bn_finalize: bad arguments");
    bn_finalize_kernel<<<(unsigned)ceil_div_l(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, mean, invstd, running_mean,
                                                                                      running_var, count, C, eps, momentum);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_bn_eval_stats(const float* running_mean, const float* running_var, float* mean, float* invstd, int C,
                                 float eps, void* stream) {
    S2S_REQUIRE(running_mean && running_var && mean && invstd && C > 0, "bn_eval_stats: bad arguments");
    bn_eval_stats_kernel<<<(unsigned)ceil_div_l(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, mean, invstd, C, eps);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_bn_apply(const void* x, const float* mean, const float* invstd, const float* gamma,
                            const float* beta, void* y, int B, int L, int halo, int C, int use_tanh,
                            const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(x && mean && invstd && gamma && beta && y && B > 0 && L > 0 && C > 0, "bn_apply: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    BNGeom g{L, L + 2 * halo, halo, C};
    Dropout d = make_dropout(drop);
    if (C % 8 == 0 && C <= 4096 && aligned16(x, y)) {
        const long nvec = (long)B * g.Lp * (C / 8);
        S2S_DISPATCH_DTYPE(dtype, T, S2S_BN_ACT_DISPATCH(use_tanh, ACT, (bn_apply_flat_kernel<T, ACT, HALO><<<bn_flat_grid(nvec), 256,
            (size_t)2 * C * sizeof(float), st>>>((const T*)x, mean, invstd, gamma, beta, (T*)y, nvec, g, d))));
        S2S_LAUNCH_OK();
        return S2S_OK;
    }
    bool ok = vec4_ok(C, C, x, y);
    long total = (long)B * g.Lp * C;
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, (bn_apply_kernel<T, VEC><<<ew_grid(total / VEC, 256), 256, 0, st>>>(
        (const T*)x, mean, invstd, gamma, beta, (T*)y, B, g, use_tanh, d))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_bn_bwd_reduce(const void* dy, const void* y, const void* x, const float* mean,
                                 const float* invstd, const float* gamma, const float* beta, float* sums, int B,
                                 int L, int halo, int C, int use_tanh, const s2s_dropout_t* drop, int dtype,
                                 void* stream) {
    S2S_REQUIRE(dy && x && mean && invstd && gamma && beta && sums, "bn_bwd_reduce: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    BNGeom g{L, L + 2 * halo, halo, C};
    Dropout d = make_dropout(drop);
    // (a flat column-owned form of this reduction measured 127-139 us vs 166 us at the C3 decoder
    // shape, but its different summation order made the multi-step graph-vs-eager comparisons of the test suite flaky at their
    // 1e-4 bound; the row-chunked reduction below stays)
    bool ok = vec4_ok(C, C, x, dy);
    int rc = S2S_OK;
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, {
        BNBwdF<T, VEC> f{(const T*)dy, (const T*)y, (const T*)x, mean, invstd, gamma, beta, g, use_tanh, d, sums};
        rc = launch_colreduce<2, VEC>(f, (long)B * L, C, st);
    }));
    return rc;
}

extern "C" int s2s_bn_bwd_apply(const void* dy, const void* y, const void* x, const float* mean,
                                const float* invstd, const float* gamma, const float* beta, const float* sums,
                                void* dx, float* dgamma, float* dbeta, int B, int L, int halo, int C, int use_tanh,
                                const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(dy && x && mean && invstd && gamma && beta && dx, "bn_bwd_apply: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    BNGeom g{L, L + 2 * halo, halo, C};
    Dropout d = make_dropout(drop);
    if (C % 8 == 0 && C <= 1536 && aligned16(x, dy, dx)) {
        const long nvec = (long)B * g.Lp * (C / 8);
        S2S_DISPATCH_DTYPE(dtype, T, S2S_BN_ACT_DISPATCH(use_tanh, ACT, (bn_bwd_apply_flat_kernel<T, ACT, HALO><<<bn_flat_grid(nvec), 256,
            (size_t)6 * C * sizeof(float), st>>>((const T*)dy, (const T*)x, mean, invstd, gamma, beta, sums, (T*)dx, nvec, g,
                                                  1.f / (float)((long)B * L), d))));
    } else {
        bool ok = vec4_ok(C, C, x, dy, dx);
        long total = (long)B * g.Lp * C;
        S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, (bn_bwd_apply_kernel<T, VEC><<<ew_grid(total / VEC, 256), 256, 0, st>>>(
            (const T*)dy, (const T*)y, (const T*)x, mean, invstd, gamma, beta, sums, (T*)dx, B, g, use_tanh, d))));
    }
    S2S_LAUNCH_OK();
    if (sums && (dgamma || dbeta)) {
        bn_param_grad_kernel<<<(unsigned)ceil_div_l(C, 128), 128, 0, st>>>(sums, dgamma, dbeta, C);
        S2S_LAUNCH_OK();
    }
    return S2S_OK;
}

static int pad_rows_impl(const void* x, void* y, int B, int L, int halo, int C, int dtype, void* stream, int to_padded) {
    S2S_REQUIRE(x && y && B > 0 && L > 0 && C > 0 && halo >= 0, "pad_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec4_ok(C, C, x, y);
    long total = (long)B * (L + 2 * halo) * C;
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, (pad_rows_kernel<T, VEC><<<ew_grid(total / VEC, 256), 256, 0, st>>>(
        (const T*)x, (T*)y, B, L, halo, C, to_padded))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_pad_rows(const void* x, void* y, int B, int L, int halo, int C, int dtype, void* stream) {
    return pad_rows_impl(x, y, B, L, halo, C, dtype, stream, 1);
}
extern "C" int s2s_unpad_rows(const void* x, void* y, int B, int L, int halo, int C, int dtype, void* stream) {
    return pad_rows_impl(x, y, B, L, halo, C, dtype, stream, 0);
}

extern "C" int s2s_skinny_linear_fwd(const void* x, const void* w, const float* bias, void* y, int64_t rows, int K, int N,
                                     int dtype, void* stream) {
    S2S_REQUIRE(x && w && y && K > 0 && N >= 1 && N <= 4, "skinny_linear_fwd: bad arguments (N must be 1..4)");
    if (rows <= 0) return S2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec4_ok(K, K, x, w);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, (skinny_fwd_kernel<T, VEC><<<(unsigned)ceil_div_l(rows, 4), 128, 0, st>>>(
        (const T*)x, (const T*)w, bias, (T*)y, rows, K, N))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_skinny_linear_bwd(const void* dy, const void* x, const void* w, float* dw, float* dbias, void* dx,
                                     int dx_accumulate, int64_t rows, int K, int N, int dtype, void* stream) {
    S2S_REQUIRE(dy && x && w && K > 0 && N >= 1 && N <= 4, "skinny_linear_bwd: bad arguments (N must be 1..4)");
    if (rows <= 0) return S2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec4_ok(K, K, x, w, dx);
    int rc = S2S_OK;
    if (dw) {
        S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, {
            SkinnyDwF<T, VEC> f{(const T*)dy, (const T*)x, K, N, dw};
            rc = launch_colreduce<4, VEC>(f, rows, K, st);
        }));
        if (rc != S2S_OK) return rc;
    }
    if (dbias) {
        rc = s2s_colsum(dy, rows, N, N, dbias, dtype, stream);
        if (rc != S2S_OK) return rc;
    }
    if (dx) {
        long total = (long)rows * K;
        S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC_DISPATCH(ok, VEC, (skinny_dx_kernel<T, VEC><<<ew_grid(total / VEC, 256), 256, 0, st>>>(
            (const T*)dy, (const T*)w, (T*)dx, rows, K, N, dx_accumulate))));
        S2S_LAUNCH_OK();
    }
    return S2S_OK;
}
