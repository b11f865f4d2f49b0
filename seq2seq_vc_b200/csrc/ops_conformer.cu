// HBM-bound kernels of the Conformer block (reference: seq2seq_vc/modules/conformer/{encoder_layer,convolution,swish}.py,
// modules/transformer/attention.py:209-305 RelPositionMultiHeadedAttention):
//   * pos_bias_u / pos_bias_v add on the query projection (fwd / bwd join),
//   * rel_shift folded into a gather-add on the score matrix (fwd) and its scatter (bwd),
//   * GLU, depthwise Conv1d over time (fwd, dx, dw), Swish(+dropout) for the FFN hidden layer,
//   * scale + up to two dropouts (RelPositionalEncoding x * sqrt(d) between two nn.Dropout's),
//   * run-sum row gather (nearest-neighbour F.interpolate over time and its adjoint).
// Activations are f32 or bf16 in HBM (channels-last), math is f32 in registers, every input element is
// read once per pass with 16-byte accesses when the channel count allows it.
#include "common.cuh"

namespace s2s {

template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]) { Vec8<T>::load(p, v); }
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]) { Vec8<T>::store(p, v); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// bf16 activations: one MUFU (tanh.approx, |err| ~ 2^-11, far below bf16's 2^-8 rounding) instead of ex2 + rcp -- the Swish / GLU
// passes over (B T, 1536) are issue-bound on the special-function unit once dropout is on
template <typename T> __device__ __forceinline__ float sigmoid_act(float x) {
    if constexpr (sizeof(T) == 2) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
        return fmaf(0.5f, t, 0.5f);
    } else {
        return sigmoidf_(x);
    }
}

// ---------------------------------------------------------------------------------------------
// q (rows, d) with row stride ldq  ->  qu = q + u, qv = q + v   (contiguous (rows, d))
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) bias_add2_kernel(const T* __restrict__ q, long ldq, const float* __restrict__ u,
                                                        const float* __restrict__ v, T* __restrict__ qu, T* __restrict__ qv,
                                                        long rows, int d) {
    const int dv = d / VEC;
    const long total = rows * dv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r;
        int c;
        if (total < 0x7fffffffL) { const unsigned iu = (unsigned)i, ru = iu / (unsigned)dv; r = ru; c = (int)(iu - ru * (unsigned)dv) * VEC; }
        else { r = i / dv; c = (int)(i - r * dv) * VEC; }
        float a[VEC], ou[VEC], ov[VEC];
        if constexpr (VEC == 8) ld8<T>(q + r * ldq + c, reinterpret_cast<float(&)[8]>(a));
        else a[0] = to_f<T>(q[r * ldq + c]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) { ou[k] = a[k] + u[c + k]; ov[k] = a[k] + v[c + k]; }
        if constexpr (VEC == 8) {
            st8<T>(qu + r * d + c, reinterpret_cast<float(&)[8]>(ou));
            st8<T>(qv + r * d + c, reinterpret_cast<float(&)[8]>(ov));
        } else {
            qu[r * d + c] = from_f<T>(ou[0]);
            qv[r * d + c] = from_f<T>(ov[0]);
        }
    }
}

// dq (strided) = a + b
template <typename T, int VEC>
__global__ void __launch_bounds__(256) add_strided_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out,
                                                          long ldo, long rows, int d) {
    const int dv = d / VEC;
    const long total = rows * dv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r;
        int c;
        if (total < 0x7fffffffL) { const unsigned iu = (unsigned)i, ru = iu / (unsigned)dv; r = ru; c = (int)(iu - ru * (unsigned)dv) * VEC; }
        else { r = i / dv; c = (int)(i - r * dv) * VEC; }
        float x[VEC], y[VEC];
        if constexpr (VEC == 8) {
            ld8<T>(a + r * d + c, reinterpret_cast<float(&)[8]>(x));
            ld8<T>(b + r * d + c, reinterpret_cast<float(&)[8]>(y));
        } else { x[0] = to_f<T>(a[r * d + c]); y[0] = to_f<T>(b[r * d + c]); }
#pragma unroll
        for (int k = 0; k < VEC; ++k) x[k] += y[k];
        if constexpr (VEC == 8) st8<T>(out + r * ldo + c, reinterpret_cast<float(&)[8]>(x));
        else out[r * ldo + c] = from_f<T>(x[0]);
    }
}

// ---------------------------------------------------------------------------------------------
// rel_shift:  S[b,h,i,j] += BD[h,b,i, T-1-i+j]  (j < T)         S: (B,H,T,ldS)   BD: (H,B,T,ldB), ldB >= 2T-1
// backward :  dBD[h,b,i,k] = dS[b,h,i, k-(T-1-i)] if 0 <= k-(T-1-i) < T else 0      (all ldB columns written)
// one warp per (b,h,i) row
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) relshift_add_kernel(T* __restrict__ S, const T* __restrict__ BD, int B, int H, int Tn,
                                                           long ldS, long ldB) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long rows = (long)B * H * Tn;
    if (row >= rows) return;
    const int i = (int)(row % Tn);
    const int h = (int)((row / Tn) % H);
    const long b = row / ((long)Tn * H);
    T* s = S + row * ldS;
    const T* bd = BD + (((long)h * B + b) * Tn + i) * ldB + (Tn - 1 - i);
    for (int j = lane; j < Tn; j += 32) s[j] = from_f<T>(to_f<T>(s[j]) + to_f<T>(bd[j]));
}

template <typename T>
__global__ void __launch_bounds__(256) relshift_bwd_kernel(const T* __restrict__ dS, T* __restrict__ dBD, int B, int H, int Tn,
                                                           long ldS, long ldB) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long rows = (long)B * H * Tn;
    if (row >= rows) return;
    const int i = (int)(row % Tn);
    const int h = (int)((row / Tn) % H);
    const long b = row / ((long)Tn * H);
    const T* s = dS + row * ldS;
    T* bd = dBD + (((long)h * B + b) * Tn + i) * ldB;
    const int off = Tn - 1 - i;
    for (int k = lane; k < (int)ldB; k += 32) {
        int j = k - off;
        bd[k] = (j >= 0 && j < Tn) ? s[j] : from_f<T>(0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// Legacy rel_shift (LegacyRelPositionMultiHeadedAttention.rel_shift, attention.py:138-157): pad one zero column on the left
// of bd (T, T), re-view the (T, T+1) block as (T+1, T), drop its first row.  In flat terms out[i][j] is element
// f = T + i T + j of the padded block: zero when f % (T+1) == 0, else bd[f / (T+1)][f % (T+1) - 1] -- rows wrap around, which is
// the behaviour checkpoints trained with it depend on.  S (B,H,T,ldS) += that; BD is (H,B,T,ldB).  One warp per (b,h,i) row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) relshift_legacy_add_kernel(T* __restrict__ S, const T* __restrict__ BD, int B, int H, int Tn,
                                                                  long ldS, long ldB) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long rows = (long)B * H * Tn;
    if (row >= rows) return;
    const int i = (int)(row % Tn);
    const int h = (int)((row / Tn) % H);
    const long b = row / ((long)Tn * H);
    T* s = S + row * ldS;
    const T* bd = BD + ((long)h * B + b) * Tn * ldB;
    for (int j = lane; j < Tn; j += 32) {
        const int f = Tn + i * Tn + j, r = f / (Tn + 1), c = f - r * (Tn + 1);
        if (c > 0) s[j] = from_f<T>(to_f<T>(s[j]) + to_f<T>(bd[(long)r * ldB + c - 1]));
    }
}

// adjoint: dBD[r][c] = dS[i][j] with r (T+1) + c + 1 = T + i T + j, zero for the T - 1 elements of bd's first row that the
// shift drops; columns >= T of the padded leading dimension are zeroed too (they are GEMM operands)
template <typename T>
__global__ void __launch_bounds__(256) relshift_legacy_bwd_kernel(const T* __restrict__ dS, T* __restrict__ dBD, int B, int H, int Tn,
                                                                  long ldS, long ldB) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long rows = (long)B * H * Tn;
    if (row >= rows) return;
    const int r = (int)(row % Tn);
    const int h = (int)((row / Tn) % H);
    const long b = row / ((long)Tn * H);
    const T* ds = dS + ((b * H + h) * (long)Tn) * ldS;
    T* bd = dBD + (((long)h * B + b) * Tn + r) * ldB;
    for (int c = lane; c < (int)ldB; c += 32) {
        float v = 0.f;
        const int q = r * (Tn + 1) + c + 1 - Tn;
        if (c < Tn && q >= 0) {
            const int i = q / Tn, j = q - i * Tn;
            v = to_f<T>(ds[(long)i * ldS + j]);
        }
        bd[c] = from_f<T>(v);
    }
}

// ---------------------------------------------------------------------------------------------
// GLU over the channel dim of (rows, 2C): y = x[:, :C] * sigmoid(x[:, C:])
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) glu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long rows, int C) {
    const int cv = C / VEC;
    const long total = rows * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r;
        int c;
        if (total < 0x7fffffffL) { const unsigned iu = (unsigned)i, ru = iu / (unsigned)cv; r = ru; c = (int)(iu - ru * (unsigned)cv) * VEC; }
        else { r = i / cv; c = (int)(i - r * cv) * VEC; }
        float a[VEC], g[VEC];
        if constexpr (VEC == 8) {
            ld8<T>(x + r * 2 * C + c, reinterpret_cast<float(&)[8]>(a));
            ld8<T>(x + r * 2 * C + C + c, reinterpret_cast<float(&)[8]>(g));
        } else { a[0] = to_f<T>(x[r * 2 * C + c]); g[0] = to_f<T>(x[r * 2 * C + C + c]); }
#pragma unroll
        for (int k = 0; k < VEC; ++k) a[k] *= sigmoid_act<T>(g[k]);
        if constexpr (VEC == 8) st8<T>(y + r * C + c, reinterpret_cast<float(&)[8]>(a));
        else y[r * C + c] = from_f<T>(a[0]);
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) glu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx,
                                                      long rows, int C) {
    const int cv = C / VEC;
    const long total = rows * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long r;
        int c;
        if (total < 0x7fffffffL) { const unsigned iu = (unsigned)i, ru = iu / (unsigned)cv; r = ru; c = (int)(iu - ru * (unsigned)cv) * VEC; }
        else { r = i / cv; c = (int)(i - r * cv) * VEC; }
        float a[VEC], g[VEC], d[VEC], da[VEC], dg[VEC];
        if constexpr (VEC == 8) {
            ld8<T>(x + r * 2 * C + c, reinterpret_cast<float(&)[8]>(a));
            ld8<T>(x + r * 2 * C + C + c, reinterpret_cast<float(&)[8]>(g));
            ld8<T>(dy + r * C + c, reinterpret_cast<float(&)[8]>(d));
        } else { a[0] = to_f<T>(x[r * 2 * C + c]); g[0] = to_f<T>(x[r * 2 * C + C + c]); d[0] = to_f<T>(dy[r * C + c]); }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float s = sigmoid_act<T>(g[k]);
            da[k] = d[k] * s;
            dg[k] = d[k] * a[k] * s * (1.f - s);
        }
        if constexpr (VEC == 8) {
            st8<T>(dx + r * 2 * C + c, reinterpret_cast<float(&)[8]>(da));
            st8<T>(dx + r * 2 * C + C + c, reinterpret_cast<float(&)[8]>(dg));
        } else { dx[r * 2 * C + c] = from_f<T>(da[0]); dx[r * 2 * C + C + c] = from_f<T>(dg[0]); }
    }
}

// ---------------------------------------------------------------------------------------------
// depthwise Conv1d over time, channels-last (B, T, C), zero padding (K-1)/2 per utterance, weights (C, K) f32.
//   fwd : y[b,t,c]  = bias[c] + sum_j w[c,j]   * x[b, t+j-pad, c]
//   dx  : dx[b,t,c] =           sum_j w[c,j]   * dy[b, t-j+pad, c]        (FLIP = true, no bias)
// Each thread owns 8 channels x TT consecutive frames; the K taps of neighbouring frames hit L1, so HBM sees every
// input element once (plus the K-1 halo rows of each tile).
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, bool FLIP>
__global__ void __launch_bounds__(128) dwconv_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, T* __restrict__ y, int B, int Tn, int C,
                                                     int K, int TT) {
    const int cv = C / VEC;
    const int tiles = (Tn + TT - 1) / TT;
    const long total = (long)B * tiles * cv;
    const int pad = (K - 1) / 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * VEC;
        const long bt = i / cv;
        const int tile = (int)(bt % tiles);
        const long b = bt / tiles;
        const int t0 = tile * TT;
        const int t1 = min(t0 + TT, Tn);
        const T* xb = x + b * (long)Tn * C + c;
        T* yb = y + b * (long)Tn * C + c;
        for (int t = t0; t < t1; ++t) {
            float acc[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = bias ? bias[c + k] : 0.f;
            for (int j = 0; j < K; ++j) {
                const int ts = FLIP ? t - j + pad : t + j - pad;
                if (ts < 0 || ts >= Tn) continue;
                float v[VEC];
                if constexpr (VEC == 8) ld8<T>(xb + (long)ts * C, reinterpret_cast<float(&)[8]>(v));
                else v[0] = to_f<T>(xb[(long)ts * C]);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] += w[(c + k) * K + j] * v[k];
            }
            if constexpr (VEC == 8) st8<T>(yb + (long)t * C, reinterpret_cast<float(&)[8]>(acc));
            else yb[(long)t * C] = from_f<T>(acc[0]);
        }
    }
}

// dw[c,j] += sum_{b,t} dy[b,t,c] * x[b, t+j-pad, c].  blockDim (32 channels, 8 row lanes); grid (C/32, row chunks).
template <typename T, int KMAX>
__global__ void __launch_bounds__(256) dwconv_dw_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ dw,
                                                        float* __restrict__ dbias, int B, int Tn, int C, int K) {
    __shared__ float red[8][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 32 + tx;
    const int pad = (K - 1) / 2;
    const long rows = (long)B * Tn;
    const long per = (rows + gridDim.y - 1) / gridDim.y;
    const long r0 = (long)blockIdx.y * per;
    const long r1 = (r0 + per < rows) ? r0 + per : rows;
    float acc[KMAX], accb = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) acc[j] = 0.f;
    if (c < C) {
        for (long r = r0 + ty; r < r1; r += 8) {
            const int t = (int)(r % Tn);
            const float g = to_f<T>(dy[r * C + c]);
            accb += g;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
                const int ts = t + j - pad;
                if (j < K && ts >= 0 && ts < Tn) acc[j] += g * to_f<T>(x[(r + j - pad) * C + c]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
        if (j >= K) break;
        __syncthreads();
        red[ty][tx] = acc[j];
        __syncthreads();
        if (ty == 0 && c < C) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) s += red[q][tx];
            atomicAdd(dw + (long)c * K + j, s);
        }
    }
    if (dbias) {
        __syncthreads();
        red[ty][tx] = accb;
        __syncthreads();
        if (ty == 0 && c < C) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) s += red[q][tx];
            atomicAdd(dbias + c, s);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Tiled depthwise conv for the common kernel sizes (K = 7 / 15 / 31): one CTA stages a (TT + K - 1) x 64-channel tile
// of the input in shared memory with 16-byte coalesced loads (raw dtype), then every thread owns ONE channel and walks
// its frames with the K weights and a K-deep sliding window in registers: 1 LDS + K FMA per output, HBM sees each
// input element once (+ (K-1)/TT halo).  FLIP selects the adjoint (dx) form.
// ---------------------------------------------------------------------------------------------
// frames per thread = K * REP (a whole number of window rotations: no partial unrolled pass), 4 thread groups per tile
template <typename T, int K> struct DwTile {
    static constexpr int TARGET = (sizeof(T) == 2) ? 32 : 16;
    static constexpr int REP = (TARGET / K) > 0 ? (TARGET / K) : 1;
    static constexpr int G = 4, FR = K * REP, TT = G * FR, ROWS = TT + K - 1;
    static constexpr size_t fwd_smem = sizeof(T) * ROWS * 64, dw_smem = sizeof(T) * (ROWS + TT) * 64;
};

template <typename T>
__device__ __forceinline__ void dw_load_tile(T* __restrict__ dst, const T* __restrict__ src_b, int t_first, int rows, int Tn, int C,
                                             int c0, bool vec) {
    constexpr int CH = 64, E = 16 / sizeof(T);          // elements per 16-byte chunk
    for (int i = threadIdx.x; i < rows * (CH / E); i += blockDim.x) {
        const int r = i / (CH / E), cg = (i % (CH / E)) * E;
        const int t = t_first + r;
        T* d = dst + r * CH + cg;
        if (t >= 0 && t < Tn && vec && c0 + cg + E <= C) {
            *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(src_b + (long)t * C + c0 + cg);
        } else {
#pragma unroll
            for (int k = 0; k < E; ++k)
                d[k] = (t >= 0 && t < Tn && c0 + cg + k < C) ? src_b[(long)t * C + c0 + cg + k] : from_f<T>(0.f);
        }
    }
}

template <typename T, int K, bool FLIP>
__global__ void __launch_bounds__(256) dwconv_tile_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, T* __restrict__ y, int Tn, int C, int vec) {
    constexpr int CH = 64, TT = DwTile<T, K>::TT, FR = DwTile<T, K>::FR, REP = DwTile<T, K>::REP, PAD = (K - 1) / 2,
                  ROWS = DwTile<T, K>::ROWS;
    __shared__ __align__(16) T xs[ROWS * CH];
    const int b = blockIdx.z, t0 = blockIdx.x * TT, c0 = blockIdx.y * CH;
    const T* xb = x + (long)b * Tn * C;
    dw_load_tile<T>(xs, xb, t0 - PAD, ROWS, Tn, C, c0, vec != 0);
    __syncthreads();
    const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
    if (c0 + c >= C) return;
    float wr[K], win[K];
#pragma unroll
    for (int j = 0; j < K; ++j) wr[j] = w[(long)(c0 + c) * K + (FLIP ? K - 1 - j : j)];
    const float bv = bias ? bias[c0 + c] : 0.f;
    const T* xp = xs + (g * FR) * CH + c;                  // next window row to fetch
#pragma unroll
    for (int j = 0; j < K - 1; ++j) { win[j] = to_f<T>(*xp); xp += CH; }
    win[K - 1] = 0.f;
    int t = t0 + g * FR;
    T* yp = y + ((long)b * Tn + t) * C + c0 + c;
    for (int rep = 0; rep < REP; ++rep) {
#pragma unroll
        for (int f = 0; f < K; ++f) {
            win[(f + K - 1) % K] = to_f<T>(*xp);
            xp += CH;
            float acc = bv;
#pragma unroll
            for (int j = 0; j < K; ++j) acc = fmaf(wr[j], win[(f + j) % K], acc);
            if (t < Tn) *yp = from_f<T>(acc);
            yp += C;
            ++t;
        }
    }
}

// dw[c,j] += sum_{b,t} dy[b,t,c] x[b,t+j-pad,c], dbias[c] += sum dy: one CTA per (64-channel tile, utterance[, time split]),
// looping over time tiles with both operands staged in shared memory; K + 1 register accumulators per thread.
template <typename T, int K>
__global__ void __launch_bounds__(256) dwconv_dw_tile_kernel(const T* __restrict__ dy, const T* __restrict__ x, float* __restrict__ dw,
                                                             float* __restrict__ dbias, int Tn, int C, int vec, int tiles_per_cta) {
    constexpr int CH = 64, TT = DwTile<T, K>::TT, G = DwTile<T, K>::G, FR = DwTile<T, K>::FR, REP = DwTile<T, K>::REP, PAD = (K - 1) / 2,
                  ROWS = DwTile<T, K>::ROWS;
    __shared__ __align__(16) T tile[(ROWS + TT) * CH];
    static_assert(sizeof(T) * (ROWS + TT) * CH >= sizeof(float) * (G - 1) * (K + 1) * CH, "reduction scratch must fit in the tiles");
    T* xs = tile;
    T* gs = tile + ROWS * CH;
    const int b = blockIdx.z, c0 = blockIdx.x * CH;
    const T* xb = x + (long)b * Tn * C;
    const T* gb = dy + (long)b * Tn * C;
    const int c = threadIdx.x & 63, g = threadIdx.x >> 6;
    float acc[K], accb = 0.f, win[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.f;
    const int tile0 = blockIdx.y * tiles_per_cta;
    for (int tl = tile0; tl < tile0 + tiles_per_cta; ++tl) {
        const int t0 = tl * TT;
        if (t0 >= Tn) break;
        __syncthreads();
        dw_load_tile<T>(xs, xb, t0 - PAD, ROWS, Tn, C, c0, vec != 0);
        dw_load_tile<T>(gs, gb, t0, TT, Tn, C, c0, vec != 0);
        __syncthreads();
        const T* xp = xs + (g * FR) * CH + c;
        const T* gp = gs + (g * FR) * CH + c;
#pragma unroll
        for (int j = 0; j < K - 1; ++j) { win[j] = to_f<T>(*xp); xp += CH; }
        win[K - 1] = 0.f;
        for (int rep = 0; rep < REP; ++rep) {
#pragma unroll
            for (int f = 0; f < K; ++f) {
                win[(f + K - 1) % K] = to_f<T>(*xp);
                xp += CH;
                const float gv = to_f<T>(*gp);              // rows past T_n are zero in the tile
                gp += CH;
                accb += gv;
#pragma unroll
                for (int j = 0; j < K; ++j) acc[j] = fmaf(gv, win[(f + j) % K], acc[j]);
            }
        }
    }
    // cross-group reduction through shared memory (re-using the x tile), then one atomic per (channel, tap)
    __syncthreads();
    float* red = reinterpret_cast<float*>(tile);                     // [(G-1)][K+1][CH] floats
    if (g > 0) {
#pragma unroll
        for (int j = 0; j < K; ++j) red[((g - 1) * (K + 1) + j) * CH + c] = acc[j];
        red[((g - 1) * (K + 1) + K) * CH + c] = accb;
    }
    __syncthreads();
    if (g == 0 && c0 + c < C) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            float v = acc[j];
#pragma unroll
            for (int q = 0; q < G - 1; ++q) v += red[(q * (K + 1) + j) * CH + c];
            atomicAdd(dw + (long)(c0 + c) * K + j, v);
        }
        if (dbias) {
            float v = accb;
#pragma unroll
            for (int q = 0; q < G - 1; ++q) v += red[(q * (K + 1) + K) * CH + c];
            atomicAdd(dbias + c0 + c, v);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Swish (+dropout):  y = dropout(x * sigmoid(x));   dx = dy * mask * (s + x s (1 - s))
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, bool BWD>
__global__ void __launch_bounds__(256) swish_kernel(const T* __restrict__ g, const T* __restrict__ x, T* __restrict__ out, long nv,
                                                    Dropout drop) {
    dropout_resolve(drop);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) {
        float a[VEC], d[VEC];
        if constexpr (VEC == 8) ld8<T>(x + i * 8, reinterpret_cast<float(&)[8]>(a));
        else a[0] = to_f<T>(x[i]);
        if (BWD) {
            if constexpr (VEC == 8) ld8<T>(g + i * 8, reinterpret_cast<float(&)[8]>(d));
            else d[0] = to_f<T>(g[i]);
        }
        float mk[8];
        if constexpr (VEC == 8) dropout_factors<8>(drop, (uint64_t)i * 8, mk);
        else mk[0] = dropout_factor(drop, (uint64_t)i);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const float s = sigmoid_act<T>(a[k]);
            const float m = mk[k];
            a[k] = BWD ? d[k] * m * (s + a[k] * s * (1.f - s)) : a[k] * s * m;
        }
        if constexpr (VEC == 8) st8<T>(out + i * 8, reinterpret_cast<float(&)[8]>(a));
        else out[i] = from_f<T>(a[0]);
    }
}

// y = x * scale * mask1 * mask2   (its own adjoint)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) scale_dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long nv, float scale,
                                                            Dropout d1, Dropout d2) {
    dropout_resolve(d1);
    dropout_resolve(d2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) {
        float a[VEC];
        if constexpr (VEC == 8) ld8<T>(x + i * 8, reinterpret_cast<float(&)[8]>(a));
        else a[0] = to_f<T>(x[i]);
        float m1[8], m2[8];
        if constexpr (VEC == 8) {
            dropout_factors<8>(d1, (uint64_t)i * 8, m1);
            dropout_factors<8>(d2, (uint64_t)i * 8, m2);
        } else {
            m1[0] = dropout_factor(d1, (uint64_t)i);
            m2[0] = dropout_factor(d2, (uint64_t)i);
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) a[k] *= scale * m1[k] * m2[k];
        if constexpr (VEC == 8) st8<T>(y + i * 8, reinterpret_cast<float(&)[8]>(a));
        else y[i] = from_f<T>(a[0]);
    }
}

// y += alpha * x
template <typename T, int VEC>
__global__ void __launch_bounds__(256) axpy_kernel(const T* __restrict__ x, T* __restrict__ y, long nv, float alpha) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long)gridDim.x * blockDim.x) {
        float a[VEC], b[VEC];
        if constexpr (VEC == 8) { ld8<T>(x + i * 8, reinterpret_cast<float(&)[8]>(a)); ld8<T>(y + i * 8, reinterpret_cast<float(&)[8]>(b)); }
        else { a[0] = to_f<T>(x[i]); b[0] = to_f<T>(y[i]); }
#pragma unroll
        for (int k = 0; k < VEC; ++k) b[k] += alpha * a[k];
        if constexpr (VEC == 8) st8<T>(y + i * 8, reinterpret_cast<float(&)[8]>(b));
        else y[i] = from_f<T>(b[0]);
    }
}

// out[r, :] = s[r] * x[r, :]      (s f32 per row)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) rowscale_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ out,
                                                       long rows, int C) {
    const int cv = C / VEC;
    const long total = rows * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / cv;
        float a[VEC];
        const float f = s[r];
        if constexpr (VEC == 8) ld8<T>(x + i * 8, reinterpret_cast<float(&)[8]>(a));
        else a[0] = to_f<T>(x[i]);
#pragma unroll
        for (int k = 0; k < VEC; ++k) a[k] *= f;
        if constexpr (VEC == 8) st8<T>(out + i * 8, reinterpret_cast<float(&)[8]>(a));
        else out[i] = from_f<T>(a[0]);
    }
}

// y[b, i, :] = sum_{j = start[i]}^{start[i] + count[i] - 1} x[b, j, :]      x: (B, Tin, C), y: (B, Tout, C)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const T* __restrict__ x, const int32_t* __restrict__ start,
                                                          const int32_t* __restrict__ count, T* __restrict__ y, int B, int Tin,
                                                          int Tout, int C) {
    const int cv = C / VEC;
    const long total = (long)B * Tout * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * VEC;
        const long bt = i / cv;
        const int t = (int)(bt % Tout);
        const long b = bt / Tout;
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
        const int s0 = start[t], n = count[t];
        for (int j = s0; j < s0 + n; ++j) {
            if (j < 0 || j >= Tin) continue;
            float v[VEC];
            if constexpr (VEC == 8) ld8<T>(x + (b * Tin + j) * C + c, reinterpret_cast<float(&)[8]>(v));
            else v[0] = to_f<T>(x[(b * Tin + j) * C + c]);
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] += v[k];
        }
        if constexpr (VEC == 8) st8<T>(y + (b * Tout + t) * C + c, reinterpret_cast<float(&)[8]>(acc));
        else y[(b * Tout + t) * C + c] = from_f<T>(acc[0]);
    }
}

}  // namespace s2s

using namespace s2s;

#define S2S_VEC8_DISPATCH(ok, VEC, ...)                 \
    do {                                                \
        if (ok) { constexpr int VEC = 8; __VA_ARGS__; } \
        else { constexpr int VEC = 1; __VA_ARGS__; }    \
    } while (0)

static inline bool vec8_ok(long cols, long ld, const void* a, const void* b = nullptr, const void* c = nullptr,
                           const void* d = nullptr) {
    return (cols % 8 == 0) && (ld % 8 == 0) && aligned16(a, b, c, d);
}

extern "C" int s2s_bias_add2(const void* q, int64_t ldq, const float* u, const float* v, void* qu, void* qv, int64_t rows,
                             int d, int dtype, void* stream) {
    S2S_REQUIRE(q && u && v && qu && qv && rows > 0 && d > 0 && ldq >= d, "bias_add2: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec8_ok(d, ldq, q, qu, qv);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (bias_add2_kernel<T, VEC><<<ew_grid(rows * d / VEC, 256), 256, 0, st>>>(
        (const T*)q, ldq, u, v, (T*)qu, (T*)qv, rows, d))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_add_strided(const void* a, const void* b, void* out, int64_t ldo, int64_t rows, int d, int dtype,
                               void* stream) {
    S2S_REQUIRE(a && b && out && rows > 0 && d > 0 && ldo >= d, "add_strided: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec8_ok(d, ldo, a, b, out);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (add_strided_kernel<T, VEC><<<ew_grid(rows * d / VEC, 256), 256, 0, st>>>(
        (const T*)a, (const T*)b, (T*)out, ldo, rows, d))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_relshift_add(void* S, const void* BD, int B, int H, int T, int64_t ldS, int64_t ldB, int dtype,
                                void* stream) {
    S2S_REQUIRE(S && BD && B > 0 && H > 0 && T > 0 && ldS >= T && ldB >= 2 * T - 1, "relshift_add: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long rows = (long)B * H * T;
    S2S_DISPATCH_DTYPE(dtype, TT, (relshift_add_kernel<TT><<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(
        (TT*)S, (const TT*)BD, B, H, T, ldS, ldB)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_relshift_bwd(const void* dS, void* dBD, int B, int H, int T, int64_t ldS, int64_t ldB, int dtype,
                                void* stream) {
    S2S_REQUIRE(dS && dBD && B > 0 && H > 0 && T > 0 && ldS >= T && ldB >= 2 * T - 1, "relshift_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long rows = (long)B * H * T;
    S2S_DISPATCH_DTYPE(dtype, TT, (relshift_bwd_kernel<TT><<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(
        (const TT*)dS, (TT*)dBD, B, H, T, ldS, ldB)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_relshift_legacy_add(void* S, const void* BD, int B, int H, int T, int64_t ldS, int64_t ldB, int dtype,
                                       void* stream) {
    S2S_REQUIRE(S && BD && B > 0 && H > 0 && T > 0 && ldS >= T && ldB >= T && T < 46000, "relshift_legacy_add: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long rows = (long)B * H * T;
    S2S_DISPATCH_DTYPE(dtype, TT, (relshift_legacy_add_kernel<TT><<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(
        (TT*)S, (const TT*)BD, B, H, T, ldS, ldB)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_relshift_legacy_bwd(const void* dS, void* dBD, int B, int H, int T, int64_t ldS, int64_t ldB, int dtype,
                                       void* stream) {
    S2S_REQUIRE(dS && dBD && B > 0 && H > 0 && T > 0 && ldS >= T && ldB >= T && T < 46000, "relshift_legacy_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long rows = (long)B * H * T;
    S2S_DISPATCH_DTYPE(dtype, TT, (relshift_legacy_bwd_kernel<TT><<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(
        (const TT*)dS, (TT*)dBD, B, H, T, ldS, ldB)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_glu_fwd(const void* x, void* y, int64_t rows, int C, int dtype, void* stream) {
    S2S_REQUIRE(x && y && rows > 0 && C > 0, "glu_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec8_ok(C, C, x, y);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (glu_fwd_kernel<T, VEC><<<ew_grid(rows * C / VEC, 256), 256, 0, st>>>(
        (const T*)x, (T*)y, rows, C))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_glu_bwd(const void* dy, const void* x, void* dx, int64_t rows, int C, int dtype, void* stream) {
    S2S_REQUIRE(dy && x && dx && rows > 0 && C > 0, "glu_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec8_ok(C, C, x, dy, dx);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (glu_bwd_kernel<T, VEC><<<ew_grid(rows * C / VEC, 256), 256, 0, st>>>(
        (const T*)dy, (const T*)x, (T*)dx, rows, C))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

template <typename TY, bool FLIP>
static bool launch_dwconv_tile(const TY* x, const float* w, const float* bias, TY* y, int B, int T, int C, int K, cudaStream_t st) {
    if (B > 65535) return false;
    const int vec = (C % (16 / (int)sizeof(TY)) == 0) && aligned16(x, y);
    const unsigned gy = (unsigned)ceil_div_l(C, 64);
#define S2S_DW_FWD(KK)                                                                                                 \
    do {                                                                                                               \
        dim3 grid((unsigned)ceil_div_l(T, DwTile<TY, KK>::TT), gy, (unsigned)B);                                       \
        dwconv_tile_kernel<TY, KK, FLIP><<<grid, 256, 0, st>>>(x, w, bias, y, T, C, vec);                               \
    } while (0)
    if (K == 7) S2S_DW_FWD(7);
    else if (K == 15) S2S_DW_FWD(15);
    else if (K == 31) S2S_DW_FWD(31);
    else return false;
#undef S2S_DW_FWD
    return true;
}

template <typename TY>
static bool launch_dwconv_dw_tile(const TY* dy, const TY* x, float* dw, float* dbias, int B, int T, int C, int K, cudaStream_t st) {
    if (B > 65535) return false;
    const int vec = (C % (16 / (int)sizeof(TY)) == 0) && aligned16(x, dy);
    const long ctas_full = ceil_div_l(C, 64) * B;
#define S2S_DW_DW(KK)                                                                                                  \
    do {                                                                                                               \
        if constexpr (DwTile<TY, KK>::dw_smem <= 48 * 1024) {                                                          \
            const int tiles = (int)ceil_div_l(T, DwTile<TY, KK>::TT);                                                  \
            int split = 1; /* split time only when (channel tiles x utterances) cannot fill the chip */                \
            while (ctas_full * split < 2L * num_sms() && split < tiles) split *= 2;                                    \
            const int per = (int)ceil_div_l(tiles, split);                                                             \
            dim3 grid((unsigned)ceil_div_l(C, 64), (unsigned)ceil_div_l(tiles, per), (unsigned)B);                     \
            dwconv_dw_tile_kernel<TY, KK><<<grid, 256, 0, st>>>(dy, x, dw, dbias, T, C, vec, per);                      \
            return true;                                                                                               \
        }                                                                                                              \
        return false;                                                                                                  \
    } while (0)
    if (K == 7) S2S_DW_DW(7);
    else if (K == 15) S2S_DW_DW(15);
    else if (K == 31) S2S_DW_DW(31);
#undef S2S_DW_DW
    return false;
}

extern "C" int s2s_dwconv_fwd(const void* x, const float* w, const float* bias, void* y, int B, int T, int C, int K,
                              int dtype, void* stream) {
    S2S_REQUIRE(x && w && y && B > 0 && T > 0 && C > 0 && K > 0 && (K & 1), "dwconv_fwd: bad arguments (odd K required)");
    cudaStream_t st = (cudaStream_t)stream;
    bool done = false;
    S2S_DISPATCH_DTYPE(dtype, TY, done = launch_dwconv_tile<TY, false>((const TY*)x, w, bias, (TY*)y, B, T, C, K, st));
    if (!done) {
        bool ok = vec8_ok(C, C, x, y);
        const int TT = 16;
        long items = (long)B * ceil_div_l(T, TT) * (C / (ok ? 8 : 1));
        S2S_DISPATCH_DTYPE(dtype, TY, S2S_VEC8_DISPATCH(ok, VEC, (dwconv_kernel<TY, VEC, false><<<ew_grid(items, 128), 128, 0, st>>>(
            (const TY*)x, w, bias, (TY*)y, B, T, C, K, TT))));
    }
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_dwconv_bwd(const void* dy, const void* x, const float* w, void* dx, float* dw, float* dbias, int B, int T,
                              int C, int K, int dtype, void* stream) {
    S2S_REQUIRE(dy && x && w && B > 0 && T > 0 && C > 0 && K > 0 && (K & 1) && K <= 63, "dwconv_bwd: bad arguments (odd K <= 63)");
    S2S_REQUIRE(dbias == nullptr || dw != nullptr, "dwconv_bwd: dbias is produced together with dw");
    cudaStream_t st = (cudaStream_t)stream;
    if (dx) {
        bool done = false;
        S2S_DISPATCH_DTYPE(dtype, TY, done = launch_dwconv_tile<TY, true>((const TY*)dy, w, nullptr, (TY*)dx, B, T, C, K, st));
        if (!done) {
            bool ok = vec8_ok(C, C, dy, dx);
            const int TT = 16;
            long items = (long)B * ceil_div_l(T, TT) * (C / (ok ? 8 : 1));
            S2S_DISPATCH_DTYPE(dtype, TY, S2S_VEC8_DISPATCH(ok, VEC, (dwconv_kernel<TY, VEC, true><<<ew_grid(items, 128), 128, 0, st>>>(
                (const TY*)dy, w, nullptr, (TY*)dx, B, T, C, K, TT))));
        }
        S2S_LAUNCH_OK();
    }
    if (dw) {
        bool done = false;
        S2S_DISPATCH_DTYPE(dtype, TY, done = launch_dwconv_dw_tile<TY>((const TY*)dy, (const TY*)x, dw, dbias, B, T, C, K, st));
        if (!done) {
            unsigned gx = (unsigned)ceil_div_l(C, 32);
            long rows = (long)B * T;
            long want = (long)num_sms() * 4 / gx;
            if (want < 1) want = 1;
            long maxy = ceil_div_l(rows, 64);
            if (want > maxy) want = maxy;
            dim3 grid(gx, (unsigned)want), block(32, 8);
            S2S_DISPATCH_DTYPE(dtype, TY, {
                if (K <= 7) dwconv_dw_kernel<TY, 7><<<grid, block, 0, st>>>((const TY*)dy, (const TY*)x, dw, dbias, B, T, C, K);
                else if (K <= 15) dwconv_dw_kernel<TY, 15><<<grid, block, 0, st>>>((const TY*)dy, (const TY*)x, dw, dbias, B, T, C, K);
                else if (K <= 31) dwconv_dw_kernel<TY, 31><<<grid, block, 0, st>>>((const TY*)dy, (const TY*)x, dw, dbias, B, T, C, K);
                else dwconv_dw_kernel<TY, 63><<<grid, block, 0, st>>>((const TY*)dy, (const TY*)x, dw, dbias, B, T, C, K);
            });
        }
        S2S_LAUNCH_OK();
    }
    return S2S_OK;
}

extern "C" int s2s_swish_fwd(const void* x, void* y, int64_t n, const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(x && y && n > 0, "swish_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    Dropout d = make_dropout(drop);
    bool ok = vec8_ok(n, 8, x, y);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (swish_kernel<T, VEC, false><<<ew_grid(n / VEC, 256), 256, 0, st>>>(
        nullptr, (const T*)x, (T*)y, n / VEC, d))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_swish_bwd(const void* dy, const void* x, void* dx, int64_t n, const s2s_dropout_t* drop, int dtype,
                             void* stream) {
    S2S_REQUIRE(dy && x && dx && n > 0, "swish_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    Dropout d = make_dropout(drop);
    bool ok = vec8_ok(n, 8, x, dy, dx);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (swish_kernel<T, VEC, true><<<ew_grid(n / VEC, 256), 256, 0, st>>>(
        (const T*)dy, (const T*)x, (T*)dx, n / VEC, d))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_scale_dropout(const void* x, void* y, int64_t n, float scale, const s2s_dropout_t* drop1,
                                 const s2s_dropout_t* drop2, int dtype, void* stream) {
    S2S_REQUIRE(x && y && n > 0, "scale_dropout: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    Dropout d1 = make_dropout(drop1), d2 = make_dropout(drop2);
    bool ok = vec8_ok(n, 8, x, y);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (scale_dropout_kernel<T, VEC><<<ew_grid(n / VEC, 256), 256, 0, st>>>(
        (const T*)x, (T*)y, n / VEC, scale, d1, d2))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_axpy(const void* x, void* y, int64_t n, float alpha, int dtype, void* stream) {
    S2S_REQUIRE(x && y && n > 0, "axpy: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec8_ok(n, 8, x, y);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (axpy_kernel<T, VEC><<<ew_grid(n / VEC, 256), 256, 0, st>>>(
        (const T*)x, (T*)y, n / VEC, alpha))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_rowscale(const void* x, const float* s, void* out, int64_t rows, int C, int dtype, void* stream) {
    S2S_REQUIRE(x && s && out && rows > 0 && C > 0, "rowscale: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec8_ok(C, C, x, out);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (rowscale_kernel<T, VEC><<<ew_grid(rows * C / VEC, 256), 256, 0, st>>>(
        (const T*)x, s, (T*)out, rows, C))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_gather_rows(const void* x, const int32_t* start, const int32_t* count, void* y, int B, int Tin, int Tout,
                               int C, int dtype, void* stream) {
    S2S_REQUIRE(x && start && count && y && B > 0 && Tin > 0 && Tout > 0 && C > 0, "gather_rows: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = vec8_ok(C, C, x, y);
    S2S_DISPATCH_DTYPE(dtype, T, S2S_VEC8_DISPATCH(ok, VEC, (gather_rows_kernel<T, VEC><<<ew_grid((long)B * Tout * C / VEC, 256), 256, 0, st>>>(
        (const T*)x, start, count, (T*)y, B, Tin, Tout, C))));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
