// Remaining HBM-bound pieces of the VTN/TransformerTTS step: scaled positional encoding,
// the Conv2dSubsampling front-end (direct 1->C conv, stride-2 im2col/col2im), decoder input and
// target glue, the Seq2Seq / guided-attention losses (value + gradient in one pass), the
// clip-norm + Adam tail over flat buffers, and dtype casts / weight packing.
#include "common.cuh"

namespace s2s {

// ---------------------------------------------------------------------------------------------
// ScaledPositionalEncoding
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void scaled_pe_fwd_kernel(const T* __restrict__ x, const float* __restrict__ pe,
                                     const float* __restrict__ alpha, T* __restrict__ y, long n, long td,
                                     Dropout drop) {
    dropout_resolve(drop);
    const float a = *alpha;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float v = to_f<T>(x[i]) + a * pe[i % td];
        y[i] = from_f<T>(v * dropout_factor(drop, (uint64_t)i));
    }
}
template <typename T>
__global__ void scaled_pe_bwd_kernel(const T* __restrict__ dy, const float* __restrict__ pe, T* __restrict__ dx,
                                     float* dalpha, long n, long td, Dropout drop) {
    __shared__ float red[32];
    dropout_resolve(drop);
    float acc = 0.f;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float g = to_f<T>(dy[i]) * dropout_factor(drop, (uint64_t)i);
        if (dx) dx[i] = from_f<T>(g);
        acc += g * pe[i % td];
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && dalpha) atomicAdd(dalpha, acc);
}

// ---------------------------------------------------------------------------------------------
// conv1: (B, T, F) -> relu(conv2d 1->C, 3x3, stride 2) channels-last (B, T1, F1, C)
// ---------------------------------------------------------------------------------------------
// Token embedding + <eos> append + ScaledPositionalEncoding (TransformerTTS encoder input layer)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long tts_token(const int64_t* __restrict__ tokens, const int32_t* __restrict__ ilens, int b, int t,
                                          int T_in, int eos, int pad) {
    const int il = ilens[b];
    if (t < il && t < T_in) return tokens[(long)b * T_in + t];
    return (t == il) ? eos : pad;
}
template <typename T>
__global__ void embed_pe_fwd_kernel(const int64_t* __restrict__ tokens, const int32_t* __restrict__ ilens,
                                    const float* __restrict__ w, const float* __restrict__ pe, const float* __restrict__ alpha,
                                    T* __restrict__ y, int B, int T_in, int T_out, int d, int eos, int pad, Dropout drop) {
    dropout_resolve(drop);
    const float a = *alpha;
    const long n = (long)B * T_out * d;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % d);
        const long q = i / d;
        const int t = (int)(q % T_out);
        const int b = (int)(q / T_out);
        const long tok = tts_token(tokens, ilens, b, t, T_in, eos, pad);
        const float v = w[tok * d + c] + a * pe[(long)t * d + c];
        y[i] = from_f<T>(v * dropout_factor(drop, (uint64_t)i));
    }
}
template <typename T>
__global__ void __launch_bounds__(256) embed_pe_bwd_kernel(const T* __restrict__ dy, const int64_t* __restrict__ tokens,
                                                           const int32_t* __restrict__ ilens, const float* __restrict__ pe,
                                                           float* __restrict__ dw, float* dalpha, int B, int T_in, int T_out, int d,
                                                           int eos, int pad, Dropout drop) {
    __shared__ float red[32];
    dropout_resolve(drop);
    float acc = 0.f;
    const long n = (long)B * T_out * d;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % d);
        const long q = i / d;
        const int t = (int)(q % T_out);
        const int b = (int)(q / T_out);
        const float g = to_f<T>(dy[i]) * dropout_factor(drop, (uint64_t)i);
        acc += g * pe[(long)t * d + c];
        const long tok = tts_token(tokens, ilens, b, t, T_in, eos, pad);
        if (dw && tok != pad) atomicAdd(dw + tok * d + c, g);       // padding_idx row never receives a gradient
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && dalpha) atomicAdd(dalpha, acc);
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void conv1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                 const float* __restrict__ bias, T* __restrict__ y, int B, int Tn, int F, int C,
                                 int T1, int F1) {
    extern __shared__ float sw[];  // [9][C] weights then [C] bias
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) sw[(i % 9) * C + i / 9] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sw[9 * C + i] = bias[i];
    __syncthreads();
    const long total = (long)B * T1 * F1 * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long pos = i / C;
        int f1 = (int)(pos % F1);
        long r = pos / F1;
        int t1 = (int)(r % T1);
        long b = r / T1;
        const float* xp = x + (b * Tn + 2 * t1) * F + 2 * f1;
        float acc = sw[9 * C + c];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt)
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) acc = fmaf(xp[kt * F + kf], sw[(kt * 3 + kf) * C + c], acc);
        y[i] = from_f<T>(fmaxf(acc, 0.f));
    }
}

// 8 output channels per thread: one index decomposition and 9 input loads feed 72 FMAs and a 16 B store
template <typename T>
__global__ void __launch_bounds__(256) conv1_fwd8_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, T* __restrict__ y, int B, int Tn, int F,
                                                         int C, int T1, int F1) {
    extern __shared__ float sw[];  // [9][C] weights then [C] bias
    for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) sw[(i % 9) * C + i / 9] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) sw[9 * C + i] = bias[i];
    __syncthreads();
    const int cg = C / 8;
    const long total = (long)B * T1 * F1 * cg;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c, f1, t1;
        long pos, b;
        if (total < 0x7fffffffL) {                      // 32-bit index split (three 64-bit divisions cost more than the 72 FMAs)
            const unsigned iu = (unsigned)i, pu = iu / (unsigned)cg, ru = pu / (unsigned)F1, bu = ru / (unsigned)T1;
            c = (int)(iu - pu * (unsigned)cg) * 8;
            f1 = (int)(pu - ru * (unsigned)F1);
            t1 = (int)(ru - bu * (unsigned)T1);
            pos = pu; b = bu;
        } else {
            c = (int)(i % cg) * 8;
            pos = i / cg;
            f1 = (int)(pos % F1);
            const long r = pos / F1;
            t1 = (int)(r % T1);
            b = r / T1;
        }
        const float* xp = x + (b * Tn + 2 * t1) * F + 2 * f1;
        float xv[9];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt)
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) xv[kt * 3 + kf] = xp[kt * F + kf];
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = sw[9 * C + c + k];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float4 w0 = *reinterpret_cast<const float4*>(sw + tap * C + c);
            const float4 w1 = *reinterpret_cast<const float4*>(sw + tap * C + c + 4);
            acc[0] = fmaf(xv[tap], w0.x, acc[0]); acc[1] = fmaf(xv[tap], w0.y, acc[1]);
            acc[2] = fmaf(xv[tap], w0.z, acc[2]); acc[3] = fmaf(xv[tap], w0.w, acc[3]);
            acc[4] = fmaf(xv[tap], w1.x, acc[4]); acc[5] = fmaf(xv[tap], w1.y, acc[5]);
            acc[6] = fmaf(xv[tap], w1.z, acc[6]); acc[7] = fmaf(xv[tap], w1.w, acc[7]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaxf(acc[k], 0.f);
        Vec8<T>::store(y + pos * C + c, acc);
    }
}


// Channel-owned variant: thread (tx, ty) keeps the 9 x 8 weights (+ bias) of ITS 8 output channels in registers and strides
// over output positions: per position 9 broadcast input loads, 72 FMAs, one 16-byte store.  (The grid-stride kernel above
// re-fetches 18 x 16 B of weights from shared memory per output vector with a 2-way bank conflict and is bound by that.)
template <typename T>
__global__ void __launch_bounds__(256) conv1_fwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, T* __restrict__ y, int B, int Tn, int F,
                                                            int C, int T1, int F1) {
    const int c = threadIdx.x * 8;
    float wr[9][8], bs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        bs[k] = bias[c + k];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) wr[tap][k] = w[(c + k) * 9 + tap];
    }
    const unsigned P = (unsigned)B * (unsigned)T1 * (unsigned)F1;
    const unsigned stride = gridDim.x * blockDim.y;
    // two output positions per iteration: 18 independent broadcast loads in flight and two FMA chains per channel
    for (unsigned pos = blockIdx.x * blockDim.y + threadIdx.y; pos < P; pos += 2 * stride) {
        const unsigned pos2 = pos + stride;
        const bool two = pos2 < P;
        const unsigned r = pos / (unsigned)F1, b = r / (unsigned)T1;
        const int f1 = (int)(pos - r * (unsigned)F1), t1 = (int)(r - b * (unsigned)T1);
        const unsigned pq = two ? pos2 : pos;
        const unsigned r2 = pq / (unsigned)F1, b2 = r2 / (unsigned)T1;
        const int f12 = (int)(pq - r2 * (unsigned)F1), t12 = (int)(r2 - b2 * (unsigned)T1);
        const float* xp = x + ((long)b * Tn + 2 * t1) * F + 2 * f1;
        const float* xq = x + ((long)b2 * Tn + 2 * t12) * F + 2 * f12;
        float xa[9], xb[9];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt)
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) { xa[kt * 3 + kf] = xp[kt * F + kf]; xb[kt * 3 + kf] = xq[kt * F + kf]; }
        float acc[8], acc2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = acc2[k] = bs[k];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int k = 0; k < 8; ++k) { acc[k] = fmaf(xa[tap], wr[tap][k], acc[k]); acc2[k] = fmaf(xb[tap], wr[tap][k], acc2[k]); }
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] = fmaxf(acc[k], 0.f); acc2[k] = fmaxf(acc2[k], 0.f); }
        Vec8<T>::store(y + (long)pos * C + c, acc);
        if (two) Vec8<T>::store(y + (long)pos2 * C + c, acc2);
    }
}

// weight gradient with 8 channels per thread: blockDim = (32, 8); x = channel group, y = position stripe
template <typename T>
__global__ void __launch_bounds__(256) conv1_bwd8_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* dw,
                                                         float* dbias, int B, int Tn, int F, int C, int T1, int F1) {
    __shared__ float red[10][256];
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int i = ty * 32 + tx; i < 10 * 256; i += 256) (&red[0][0])[i] = 0.f;
    __syncthreads();
    const int c0 = (blockIdx.x * 32 + tx) * 8;
    const long rows = (long)B * T1 * F1;
    const long per = (rows + gridDim.y - 1) / gridDim.y;
    const long r0 = (long)blockIdx.y * per;
    const long r1 = (r0 + per < rows) ? r0 + per : rows;
    float acc[10][8];
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
    if (c0 < C) {
        // four positions per iteration: their dy vectors (kept packed) are requested before any of them is consumed -- with one
        // 16-byte load in flight per thread and 16 warps per SM the kernel ran at 1.2 TB/s (latency-bound: 207 us at the C2 shape)
        constexpr int U = 4;
        for (long rb = r0 + ty; rb < r1; rb += 8 * U) {
            Raw8<T> raw[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long r = rb + 8 * u;
                if (r < r1) raw[u].load(dy + r * C + c0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long r = rb + 8 * u;
                if (r >= r1) break;
                int f1, t1;
                long b;
                if (rows < 0x7fffffffL) {
                    const unsigned ru = (unsigned)r, qu = ru / (unsigned)F1, bu = qu / (unsigned)T1;
                    f1 = (int)(ru - qu * (unsigned)F1);
                    t1 = (int)(qu - bu * (unsigned)T1);
                    b = bu;
                } else {
                    f1 = (int)(r % F1);
                    const long q = r / F1;
                    t1 = (int)(q % T1);
                    b = q / T1;
                }
                const float* xp = x + (b * Tn + 2 * t1) * F + 2 * f1;
                float g[8];
                raw[u].get(g);
#pragma unroll
                for (int kt = 0; kt < 3; ++kt)
#pragma unroll
                    for (int kf = 0; kf < 3; ++kf) {
                        const float xv = xp[kt * F + kf];
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[kt * 3 + kf][j] = fmaf(g[j], xv, acc[kt * 3 + kf][j]);
                    }
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[9][j] += g[j];
            }
        }
#pragma unroll
        for (int k = 0; k < 10; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(&red[k][tx * 8 + j], acc[k][j]);
    }
    __syncthreads();
    for (int e = ty * 32 + tx; e < 10 * 256; e += 256) {
        const int k = e / 256, cc = e % 256;
        const int ch = blockIdx.x * 256 + cc;
        if (ch < C) {
            const float v = red[k][cc];
            if (k < 9) { if (dw) atomicAdd(dw + ch * 9 + k, v); }
            else if (dbias) atomicAdd(dbias + ch, v);
        }
    }
}

// dw[c][tap] += sum_pos dy[pos][c] * x[pos @ tap]; dbias[c] += sum_pos dy[pos][c]
// blockDim = (32, 8): x = channel, y = position stripe.
template <typename T>
__global__ void __launch_bounds__(256) conv1_bwd_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                                        float* dw, float* dbias, int B, int Tn, int F, int C, int T1,
                                                        int F1) {
    __shared__ float red[8][10][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.x * 32 + tx;
    const long rows = (long)B * T1 * F1;
    const long per = (rows + gridDim.y - 1) / gridDim.y;
    const long r0 = (long)blockIdx.y * per;
    const long r1 = (r0 + per < rows) ? r0 + per : rows;
    float acc[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[k] = 0.f;
    if (c < C) {
        for (long r = r0 + ty; r < r1; r += 8) {
            int f1 = (int)(r % F1);
            long q = r / F1;
            int t1 = (int)(q % T1);
            long b = q / T1;
            const float* xp = x + (b * Tn + 2 * t1) * F + 2 * f1;
            float g = to_f<T>(dy[r * C + c]);
#pragma unroll
            for (int kt = 0; kt < 3; ++kt)
#pragma unroll
                for (int kf = 0; kf < 3; ++kf) acc[kt * 3 + kf] = fmaf(g, xp[kt * F + kf], acc[kt * 3 + kf]);
            acc[9] += g;
        }
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) red[ty][k][tx] = acc[k];
    __syncthreads();
    for (int e = ty * 32 + tx; e < 10 * 32; e += 256) {
        int k = e / 32, cc = e % 32;
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += red[wv][k][cc];
        int ch = blockIdx.x * 32 + cc;
        if (ch < C) {
            if (k < 9) { if (dw) atomicAdd(dw + ch * 9 + k, s); }
            else if (dbias) atomicAdd(dbias + ch, s);
        }
    }
}

template <typename T, int VEC>
__global__ void im2col_s2_kernel(const T* __restrict__ y1, T* __restrict__ col, int B, int T1, int F1, int C, int T2,
                                 int F2) {
    const int cv = C / VEC;
    const long total = (long)B * T2 * F2 * 9 * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c, tap, f2, t2;
        long m, b;
        if (total < 0x7fffffffL) {                      // 32-bit index split: the copy is otherwise bound by 64-bit divisions
            const unsigned iu = (unsigned)i, qu = iu / (unsigned)cv, mu = qu / 9u, ru = mu / (unsigned)F2, bu = ru / (unsigned)T2;
            c = (int)(iu - qu * (unsigned)cv) * VEC;
            tap = (int)(qu - mu * 9u);
            f2 = (int)(mu - ru * (unsigned)F2);
            t2 = (int)(ru - bu * (unsigned)T2);
            m = mu; b = bu;
        } else {
            c = (int)(i % cv) * VEC;
            const long q = i / cv;
            tap = (int)(q % 9);
            m = q / 9;
            f2 = (int)(m % F2);
            const long r = m / F2;
            t2 = (int)(r % T2);
            b = r / T2;
        }
        int kt = tap / 3, kf = tap % 3;
        const T* src = y1 + (((b * T1 + 2 * t2 + kt) * F1) + 2 * f2 + kf) * C + c;
        T* dst = col + (m * 9 + tap) * C + c;
        if (VEC == 8) {
            float v[8];
            Vec8<T>::load(src, v);
            Vec8<T>::store(dst, v);
        } else if (VEC == 4) {
            float v[4];
            Vec4<T>::load(src, v);
            Vec4<T>::store(dst, v);
        } else {
            *dst = *src;
        }
    }
}

template <typename T, int VEC>
__global__ void col2im_s2_kernel(const T* __restrict__ dcol, T* __restrict__ dy1, int B, int T1, int F1, int C,
                                 int T2, int F2, const T* __restrict__ gate) {
    const int cv = C / VEC;
    const long total = (long)B * T1 * F1 * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c, f1, t1;
        long q, b;
        if (total < 0x7fffffffL) {
            const unsigned iu = (unsigned)i, qu = iu / (unsigned)cv, ru = qu / (unsigned)F1, bu = ru / (unsigned)T1;
            c = (int)(iu - qu * (unsigned)cv) * VEC;
            f1 = (int)(qu - ru * (unsigned)F1);
            t1 = (int)(ru - bu * (unsigned)T1);
            q = qu; b = bu;
        } else {
            c = (int)(i % cv) * VEC;
            q = i / cv;
            f1 = (int)(q % F1);
            const long r = q / F1;
            t1 = (int)(r % T1);
            b = r / T1;
        }
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            int tt = t1 - kt;
            if (tt < 0 || (tt & 1)) continue;
            int t2 = tt >> 1;
            if (t2 >= T2) continue;
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
                int ff = f1 - kf;
                if (ff < 0 || (ff & 1)) continue;
                int f2 = ff >> 1;
                if (f2 >= F2) continue;
                long m = (b * T2 + t2) * F2 + f2;
                const T* src = dcol + (m * 9 + kt * 3 + kf) * C + c;
                if (VEC == 8) {
                    float v[8];
                    Vec8<T>::load(src, v);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k < VEC ? k : 0] += v[k];
                } else if (VEC == 4) {
                    float v[4];
                    Vec4<T>::load(src, v);
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[k < VEC ? k : 0] += v[k];
                } else {
                    acc[0] += to_f<T>(*src);
                }
            }
        }
        if (gate) {                                   // ReLU' of the conv1 output fused into the scatter-add: dy1 *= (y1 > 0)
            const T* gp = gate + q * C + c;
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (!(to_f<T>(gp[k]) > 0.f)) acc[k] = 0.f;
        }
        T* dst = dy1 + q * C + c;
        if (VEC == 8) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = acc[k < VEC ? k : 0];
            Vec8<T>::store(dst, o);
        } else if (VEC == 4) {
            float o[4] = {acc[0], acc[VEC > 1 ? 1 : 0], acc[VEC > 2 ? 2 : 0], acc[VEC > 3 ? 3 : 0]};
            Vec4<T>::store(dst, o);
        } else {
            *dst = from_f<T>(acc[0]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// decoder input / target glue
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void shift_thin_kernel(const float* __restrict__ ys, T* __restrict__ out, int B, int L, int Lr, int odim,
                                  int r) {
    const long total = (long)B * Lr * odim;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int c = (int)(i % odim);
        long q = i / odim;
        int l = (int)(q % Lr);
        long b = q / Lr;
        float v = 0.f;
        if (l > 0) {
            int src = l * r - 1;
            if (src < L) v = ys[(b * L + src) * odim + c];
        }
        out[i] = from_f<T>(v);
    }
}

__global__ void fix_targets_kernel(const float* __restrict__ labels, const int32_t* __restrict__ olens,
                                   float* __restrict__ labels_out, int32_t* __restrict__ olens_out, int B, int Lin,
                                   int Lout, int r) {
    const long total = (long)B * Lout;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int l = (int)(i % Lout);
        int b = (int)(i / Lout);
        int o = olens[b];
        o -= o % r;
        float v = labels[(long)b * Lin + l];
        if (l == o - 1) v = 1.f;
        labels_out[i] = v;
        if (l == 0 && olens_out) olens_out[b] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// Seq2SeqLoss: value + gradient
// ---------------------------------------------------------------------------------------------
template <typename T, typename TL>
__global__ void __launch_bounds__(256) seq2seq_loss_kernel(const T* __restrict__ after, const T* __restrict__ before,
                                                           const TL* __restrict__ logits, const float* __restrict__ ys,
                                                           const float* __restrict__ labels,
                                                           const int32_t* __restrict__ olens, int B, int L, int L_ys,
                                                           int L_lab, int odim, float pos_weight,
                                                           T* __restrict__ d_after,
                                                           T* __restrict__ d_before, TL* __restrict__ d_logits,
                                                           float* ws) {
    __shared__ float red[32];
    __shared__ float s_nf;
    if (threadIdx.x == 0) {
        long nf = 0;
        for (int b = 0; b < B; ++b) { int o = olens[b]; nf += (o < L ? (o > 0 ? o : 0) : L); }
        s_nf = (float)nf;
    }
    __syncthreads();
    const float nf = s_nf;
    const float inv_l1 = nf > 0.f ? 1.f / (nf * (float)odim) : 0.f;
    const float inv_bce = nf > 0.f ? 1.f / nf : 0.f;
    float s_after = 0.f, s_before = 0.f, s_bce = 0.f;
    const long n1 = (long)B * L * odim;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += (long)gridDim.x * blockDim.x) {
        long q = i / odim;
        int l = (int)(q % L);
        int b = (int)(q / L);
        float ga = 0.f, gb = 0.f;
        if (l < olens[b]) {
            float y = ys[((long)b * L_ys + l) * odim + (i - q * odim)];
            float da = to_f<T>(after[i]) - y, db = to_f<T>(before[i]) - y;
            s_after += fabsf(da);
            s_before += fabsf(db);
            ga = (da > 0.f ? 1.f : (da < 0.f ? -1.f : 0.f)) * inv_l1;
            gb = (db > 0.f ? 1.f : (db < 0.f ? -1.f : 0.f)) * inv_l1;
        }
        if (d_after) d_after[i] = from_f<T>(ga);
        if (d_before) d_before[i] = from_f<T>(gb);
    }
    const long n2 = (long)B * L;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) {
        int l = (int)(i % L);
        int b = (int)(i / L);
        float g = 0.f;
        if (l < olens[b]) {
            float x = to_f<TL>(logits[i]), y = labels[(long)b * L_lab + l];
            float lw = 1.f + (pos_weight - 1.f) * y;
            float sp = log1pf(__expf(-fabsf(x))) + fmaxf(-x, 0.f);  // softplus(-x)
            s_bce += (1.f - y) * x + lw * sp;
            float sig_neg = 1.f / (1.f + __expf(x));  // sigmoid(-x)
            g = ((1.f - y) - lw * sig_neg) * inv_bce;
        }
        if (d_logits) d_logits[i] = from_f<TL>(g);
    }
    s_after = block_sum(s_after, red);
    s_before = block_sum(s_before, red);
    s_bce = block_sum(s_bce, red);
    if (threadIdx.x == 0) {
        atomicAdd(ws + 0, s_after * inv_l1);
        atomicAdd(ws + 1, s_before * inv_l1);
        atomicAdd(ws + 2, s_bce * inv_bce);
    }
}
__global__ void seq2seq_loss_final_kernel(const float* ws, float* losses) {
    losses[0] = ws[0] + ws[1];
    losses[1] = ws[2];
}

// ---------------------------------------------------------------------------------------------
// Guided attention loss
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) guided_attn_kernel(const T* __restrict__ att, const int32_t* __restrict__ ilens,
                                                          const int32_t* __restrict__ olens, int B, int H, int T_out,
                                                          int T_in, long ld, float inv2s2, float alpha,
                                                          T* __restrict__ d_att, float* ws) {
    __shared__ float red[32];
    __shared__ float s_cnt;
    if (threadIdx.x == 0) {
        double cnt = 0;
        for (int b = 0; b < B; ++b) {
            int il = min(max(ilens[b], 0), T_in), ol = min(max(olens[b], 0), T_out);
            cnt += (double)il * ol * H;
        }
        s_cnt = (float)cnt;
    }
    __syncthreads();
    const float inv = s_cnt > 0.f ? alpha / s_cnt : 0.f;
    float acc = 0.f;
    const long total = (long)B * H * T_out * ld;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int s = (int)(i % ld);
        long q = i / ld;
        int t = (int)(q % T_out);
        int b = (int)(q / ((long)T_out * H));
        int il = ilens[b], ol = olens[b];
        float g = 0.f;
        if (s < il && s < T_in && t < ol) {
            float dlt = (float)s / (float)il - (float)t / (float)ol;
            float w = 1.f - __expf(-dlt * dlt * inv2s2);
            acc += w * to_f<T>(att[i]);
            g = w * inv;
        }
        if (d_att) d_att[i] = from_f<T>(g);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(ws, acc * inv);
}
__global__ void copy_scalar_kernel(const float* src, float* dst) { *dst = *src; }

// ---------------------------------------------------------------------------------------------
// optimizer tail
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ g, long n, float* out) {
    __shared__ float red[32];
    float acc = 0.f;
    const long n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 v = g4[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (long i = n4 * 4 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        acc += g[i] * g[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   bf16* __restrict__ p16, long n, const float* lr_dev, float beta1,
                                                   float beta2, float eps, float wd, const float* step_dev,
                                                   const float* sqnorm, float max_norm, float grad_scale) {
    const float lr = *lr_dev;
    const float step = *step_dev;
    float coef = grad_scale;
    if (sqnorm && max_norm > 0.f) {
        float total = sqrtf(*sqnorm) * grad_scale;
        float c = max_norm / (total + 1e-6f);
        coef *= fminf(c, 1.f);
    }
    const float bc1 = 1.f - powf(beta1, step);
    const float bc2 = 1.f - powf(beta2, step);
    const float step_size = lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float pi = p[i];
        float gi = g[i] * coef + wd * pi;
        float mi = beta1 * m[i] + (1.f - beta1) * gi;
        float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        pi -= step_size * mi / denom;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi;
        if (p16) p16[i] = __float2bfloat16_rn(pi);
    }
}
__global__ void step_advance_kernel(float* step, uint64_t* seed) {
    if (step) *step += 1.f;
    if (seed) *seed += 1ull;
}

// ---------------------------------------------------------------------------------------------
// casts / packing
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = from_f<TO>(to_f<TI>(in[i]));
}
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = from_f<T>(to_f<T>(a[i]) + to_f<T>(b[i]));
}
template <typename T>
__global__ void add4_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long n4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float x[4], y[4];
        Vec4<T>::load(a + 4 * i, x);
        Vec4<T>::load(b + 4 * i, y);
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] += y[k];
        Vec4<T>::store(out + 4 * i, x);
    }
}
// out[n][b][a] (+)= in[n][a][b]; 32x32 smem tile transpose
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) transpose_last2_kernel(const TI* __restrict__ in, TO* __restrict__ out, int A,
                                                              int Bd, int accumulate) {
    __shared__ float tile[32][33];
    const long n = blockIdx.z;
    const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
    const TI* src = in + n * (long)A * Bd;
    TO* dst = out + n * (long)A * Bd;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int a = a0 + r, b = b0 + threadIdx.x;
        if (a < A && b < Bd) tile[r][threadIdx.x] = to_f<TI>(src[(long)a * Bd + b]);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int b = b0 + r, a = a0 + threadIdx.x;
        if (a < A && b < Bd) {
            TO* d = dst + (long)b * A + a;
            float v = tile[threadIdx.x][r];
            if (accumulate) v += to_f<TO>(*d);
            *d = from_f<TO>(v);
        }
    }
}

template <typename TO>
__global__ void pack_conv1d_w_kernel(const float* __restrict__ w, TO* __restrict__ wp, TO* __restrict__ wpt, int OC,
                                     int IC, int K) {
    const long total = (long)OC * IC * K;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int k = (int)(i % K);
        long q = i / K;
        int ic = (int)(q % IC);
        int oc = (int)(q / IC);
        float v = w[i];
        if (wp) wp[((long)oc * K + k) * IC + ic] = from_f<TO>(v);
        if (wpt) wpt[((long)ic * K + (K - 1 - k)) * OC + oc] = from_f<TO>(v);
    }
}

}  // namespace s2s

using namespace s2s;

extern "C" int s2s_scaled_pe_fwd(const void* x, const float* pe, const float* alpha, void* y, int B, int T, int d,
                                 const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(x && pe && alpha && y && B > 0 && T > 0 && d > 0, "scaled_pe_fwd: bad arguments");
    long n = (long)B * T * d;
    Dropout dr = make_dropout(drop);
    S2S_DISPATCH_DTYPE(dtype, TT, (scaled_pe_fwd_kernel<TT><<<ew_grid(n, 1024), 256, 0, (cudaStream_t)stream>>>(
        (const TT*)x, pe, alpha, (TT*)y, n, (long)T * d, dr)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_scaled_pe_bwd(const void* dy, const float* pe, void* dx, float* dalpha, int B, int T, int d,
                                 const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(dy && pe && B > 0 && T > 0 && d > 0, "scaled_pe_bwd: bad arguments");
    long n = (long)B * T * d;
    Dropout dr = make_dropout(drop);
    S2S_DISPATCH_DTYPE(dtype, TT, (scaled_pe_bwd_kernel<TT><<<ew_grid(n, 2048), 256, 0, (cudaStream_t)stream>>>(
        (const TT*)dy, pe, (TT*)dx, dalpha, n, (long)T * d, dr)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_conv1_fwd(const float* x, const float* w, const float* bias, void* y1, int B, int T, int F, int C,
                             int dtype, void* stream) {
    S2S_REQUIRE(x && w && bias && y1 && B > 0 && T >= 3 && F >= 3 && C > 0, "conv1_fwd: bad arguments");
    int T1 = (T - 1) / 2, F1 = (F - 1) / 2;
    long total = (long)B * T1 * F1 * C;
    size_t smem = (size_t)10 * C * sizeof(float);
    S2S_REQUIRE(smem <= 48 * 1024, "conv1_fwd: C too large (%d)", C);
    if (C % 8 == 0 && C / 8 <= 256 && aligned16(y1) && (long)B * T1 * F1 < 0x7fffffffL) {
        const int cg = C / 8;
        int ty = 256 / cg;
        if (ty < 1) ty = 1;
        long gx = ceil_div_l((long)B * T1 * F1, (long)ty * 16);
        const long cap = (long)num_sms() * 8;
        if (gx > cap) gx = cap;
        S2S_DISPATCH_DTYPE(dtype, TT, (conv1_fwd_reg_kernel<TT><<<(unsigned)gx, dim3(cg, ty), 0, (cudaStream_t)stream>>>(
            x, w, bias, (TT*)y1, B, T, F, C, T1, F1)));
    } else if (C % 8 == 0 && aligned16(y1)) {
        S2S_DISPATCH_DTYPE(dtype, TT, (conv1_fwd8_kernel<TT><<<ew_grid(total / 8, 256), 256, smem, (cudaStream_t)stream>>>(
            x, w, bias, (TT*)y1, B, T, F, C, T1, F1)));
    } else {
        S2S_DISPATCH_DTYPE(dtype, TT, (conv1_fwd_kernel<TT><<<ew_grid(total, 1024), 256, smem, (cudaStream_t)stream>>>(
            x, w, bias, (TT*)y1, B, T, F, C, T1, F1)));
    }
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_conv1_bwd(const float* x, const void* dy1, float* dw, float* dbias, int B, int T, int F, int C,
                             int dtype, void* stream) {
    S2S_REQUIRE(x && dy1 && B > 0 && T >= 3 && F >= 3 && C > 0, "conv1_bwd: bad arguments");
    int T1 = (T - 1) / 2, F1 = (F - 1) / 2;
    long rows = (long)B * T1 * F1;
    const bool v8 = (C % 8 == 0) && aligned16(dy1);
    unsigned gx = (unsigned)ceil_div_l(C, v8 ? 256 : 32);
    long gy = (long)num_sms() * (v8 ? 2 : 4) / gx;
    if (gy < 1) gy = 1;
    if (gy > ceil_div_l(rows, 64)) gy = ceil_div_l(rows, 64);
    if (gy < 1) gy = 1;
    if (v8) {
        S2S_DISPATCH_DTYPE(dtype, TT, (conv1_bwd8_kernel<TT><<<dim3(gx, (unsigned)gy), dim3(32, 8), 0, (cudaStream_t)stream>>>(
            x, (const TT*)dy1, dw, dbias, B, T, F, C, T1, F1)));
    } else {
        S2S_DISPATCH_DTYPE(dtype, TT, (conv1_bwd_kernel<TT><<<dim3(gx, (unsigned)gy), dim3(32, 8), 0, (cudaStream_t)stream>>>(
            x, (const TT*)dy1, dw, dbias, B, T, F, C, T1, F1)));
    }
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_im2col_s2(const void* y1, void* col, int B, int T1, int F1, int C, int dtype, void* stream) {
    S2S_REQUIRE(y1 && col && B > 0 && T1 >= 3 && F1 >= 3 && C > 0, "im2col_s2: bad arguments");
    int T2 = (T1 - 1) / 2, F2 = (F1 - 1) / 2;
    long total = (long)B * T2 * F2 * 9 * C;
    bool ok = (C % 4 == 0) && ((uintptr_t)y1 % 16 == 0) && ((uintptr_t)col % 16 == 0);
    S2S_DISPATCH_DTYPE(dtype, TT, {
        if (ok && C % 8 == 0) im2col_s2_kernel<TT, 8><<<ew_grid(total / 8, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)y1, (TT*)col, B, T1, F1, C, T2, F2);
        else if (ok) im2col_s2_kernel<TT, 4><<<ew_grid(total / 4, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)y1, (TT*)col, B, T1, F1, C, T2, F2);
        else im2col_s2_kernel<TT, 1><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)y1, (TT*)col, B, T1, F1, C, T2, F2);
    });
    S2S_LAUNCH_OK();
    return S2S_OK;
}
static int col2im_s2_impl(const void* dcol, const void* y1, void* dy1, int B, int T1, int F1, int C, int dtype, void* stream) {
    S2S_REQUIRE(dcol && dy1 && B > 0 && T1 >= 3 && F1 >= 3 && C > 0, "col2im_s2: bad arguments");
    int T2 = (T1 - 1) / 2, F2 = (F1 - 1) / 2;
    long total = (long)B * T1 * F1 * C;
    bool ok = (C % 4 == 0) && ((uintptr_t)dy1 % 16 == 0) && ((uintptr_t)dcol % 16 == 0);
    S2S_DISPATCH_DTYPE(dtype, TT, {
        if (ok && C % 8 == 0) col2im_s2_kernel<TT, 8><<<ew_grid(total / 8, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)dcol, (TT*)dy1, B, T1, F1, C, T2, F2, (const TT*)y1);
        else if (ok) col2im_s2_kernel<TT, 4><<<ew_grid(total / 4, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)dcol, (TT*)dy1, B, T1, F1, C, T2, F2, (const TT*)y1);
        else col2im_s2_kernel<TT, 1><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)dcol, (TT*)dy1, B, T1, F1, C, T2, F2, (const TT*)y1);
    });
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_col2im_s2(const void* dcol, void* dy1, int B, int T1, int F1, int C, int dtype, void* stream) {
    return col2im_s2_impl(dcol, nullptr, dy1, B, T1, F1, C, dtype, stream);
}
extern "C" int s2s_col2im_s2_relu(const void* dcol, const void* y1, void* dy1, int B, int T1, int F1, int C, int dtype, void* stream) {
    S2S_REQUIRE(y1, "col2im_s2_relu: null y1");
    return col2im_s2_impl(dcol, y1, dy1, B, T1, F1, C, dtype, stream);
}

extern "C" int s2s_shift_thin(const float* ys, void* out, int B, int L, int Lr, int odim, int r, int dtype,
                              void* stream) {
    S2S_REQUIRE(ys && out && B > 0 && L > 0 && Lr > 0 && odim > 0 && r >= 1, "shift_thin: bad arguments");
    long total = (long)B * Lr * odim;
    S2S_DISPATCH_DTYPE(dtype, TT, (shift_thin_kernel<TT><<<ew_grid(total, 1024), 256, 0, (cudaStream_t)stream>>>(
        ys, (TT*)out, B, L, Lr, odim, r)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_fix_targets(const float* labels, const int32_t* olens, float* labels_out, int32_t* olens_out,
                               int B, int Lin, int Lout, int r, void* stream) {
    S2S_REQUIRE(labels && olens && labels_out && B > 0 && Lout > 0 && Lin >= Lout && r >= 1, "fix_targets: bad arguments");
    fix_targets_kernel<<<ew_grid((long)B * Lout, 256), 256, 0, (cudaStream_t)stream>>>(labels, olens, labels_out, olens_out,
                                                                                     B, Lin, Lout, r);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_seq2seq_loss(const void* after, const void* before, const void* logits, const float* ys,
                                const float* labels, const int32_t* olens, int B, int L, int L_ys, int L_lab, int odim,
                                float pos_weight, float* losses, void* d_after, void* d_before, void* d_logits,
                                float* workspace, int dtype, void* stream) {
    S2S_REQUIRE(after && before && logits && ys && labels && olens && losses && workspace && B > 0 && L > 0 && odim > 0 &&
                    L_ys >= L && L_lab >= L,
                "seq2seq_loss: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    S2S_CUDA_OK(cudaMemsetAsync(workspace, 0, 4 * sizeof(float), st));
    long n = (long)B * L * odim;
    S2S_DISPATCH_DTYPE(dtype, TT, (seq2seq_loss_kernel<TT, TT><<<ew_grid(n, 1024), 256, 0, st>>>(
        (const TT*)after, (const TT*)before, (const TT*)logits, ys, labels, olens, B, L, L_ys, L_lab, odim, pos_weight, (TT*)d_after,
        (TT*)d_before, (TT*)d_logits, workspace)));
    S2S_LAUNCH_OK();
    seq2seq_loss_final_kernel<<<1, 1, 0, st>>>(workspace, losses);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_guided_attn_loss(const void* att, const int32_t* ilens, const int32_t* olens, int B, int H,
                                    int T_out, int T_in, int64_t ld, float sigma, float alpha, float* loss,
                                    void* d_att, float* workspace, int dtype, void* stream) {
    S2S_REQUIRE(att && ilens && olens && loss && workspace && B > 0 && H > 0 && T_out > 0 && T_in > 0 && ld >= T_in && sigma > 0.f,
                "guided_attn_loss: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    S2S_CUDA_OK(cudaMemsetAsync(workspace, 0, 2 * sizeof(float), st));
    long n = (long)B * H * T_out * ld;
    S2S_DISPATCH_DTYPE(dtype, TT, (guided_attn_kernel<TT><<<ew_grid(n, 1024), 256, 0, st>>>(
        (const TT*)att, ilens, olens, B, H, T_out, T_in, (long)ld, 1.f / (2.f * sigma * sigma), alpha, (TT*)d_att, workspace)));
    S2S_LAUNCH_OK();
    copy_scalar_kernel<<<1, 1, 0, st>>>(workspace, loss);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_sqnorm(const float* g, int64_t n, float* out, void* stream) {
    S2S_REQUIRE(g && out && ((uintptr_t)g % 16 == 0), "sqnorm: bad arguments");
    if (n <= 0) return S2S_OK;
    sqnorm_kernel<<<ew_grid(n, 4096), 256, 0, (cudaStream_t)stream>>>(g, n, out);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const float* lr_dev,
                             float beta1, float beta2, float eps, float weight_decay, const float* step_dev,
                             const float* sqnorm, float max_norm, float grad_scale, void* stream) {
    S2S_REQUIRE(p && g && m && v && lr_dev && step_dev, "adam_step: null pointer");
    if (n <= 0) return S2S_OK;
    adam_kernel<<<ew_grid(n, 1024), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)p_bf16, n, lr_dev, beta1, beta2, eps,
                                                                     weight_decay, step_dev, sqnorm, max_norm, grad_scale);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_step_advance(float* step, uint64_t* seed, void* stream) {
    step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step, seed);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_cast(const void* in, void* out, int64_t n, int in_dtype, int out_dtype, void* stream) {
    S2S_REQUIRE(in && out, "cast: null pointer");
    if (n <= 0) return S2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = ew_grid(n, 1024);
    if (in_dtype == S2S_F32 && out_dtype == S2S_BF16) cast_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)in, (bf16*)out, n);
    else if (in_dtype == S2S_BF16 && out_dtype == S2S_F32) cast_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)in, (float*)out, n);
    else if (in_dtype == S2S_F32 && out_dtype == S2S_F32) cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, n);
    else if (in_dtype == S2S_BF16 && out_dtype == S2S_BF16) cast_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)in, (bf16*)out, n);
    else return set_error(S2S_ERR_INVALID, "cast: bad dtypes %d -> %d", in_dtype, out_dtype);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
    S2S_REQUIRE(a && b && out, "add: null pointer");
    if (n <= 0) return S2S_OK;
    cudaStream_t st = (cudaStream_t)stream;
    bool ok = (n % 4 == 0) && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)out % 16 == 0);
    S2S_DISPATCH_DTYPE(dtype, TT, {
        if (ok) add4_kernel<TT><<<ew_grid(n / 4, 256), 256, 0, st>>>((const TT*)a, (const TT*)b, (TT*)out, n / 4);
        else add_kernel<TT><<<ew_grid(n, 1024), 256, 0, st>>>((const TT*)a, (const TT*)b, (TT*)out, n);
    });
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_pack_conv1d_w(const float* w, void* wp, void* wpt, int OC, int IC, int K, int out_dtype, void* stream) {
    S2S_REQUIRE(w && (wp || wpt) && OC > 0 && IC > 0 && K > 0, "pack_conv1d_w: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long n = (long)OC * IC * K;
    S2S_DISPATCH_DTYPE(out_dtype, TT, (pack_conv1d_w_kernel<TT><<<ew_grid(n, 256), 256, 0, st>>>(w, (TT*)wp, (TT*)wpt, OC, IC, K)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_transpose_last2(const void* in, void* out, int N, int A, int Bd, int in_dtype, int out_dtype,
                                   int accumulate, void* stream) {
    S2S_REQUIRE(in && out && N > 0 && A > 0 && Bd > 0 && N <= 65535, "transpose_last2: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div_l(Bd, 32), (unsigned)ceil_div_l(A, 32), (unsigned)N), block(32, 8);
    if (in_dtype == S2S_F32 && out_dtype == S2S_BF16) transpose_last2_kernel<float, bf16><<<grid, block, 0, st>>>((const float*)in, (bf16*)out, A, Bd, accumulate);
    else if (in_dtype == S2S_F32 && out_dtype == S2S_F32) transpose_last2_kernel<float, float><<<grid, block, 0, st>>>((const float*)in, (float*)out, A, Bd, accumulate);
    else if (in_dtype == S2S_BF16 && out_dtype == S2S_F32) transpose_last2_kernel<bf16, float><<<grid, block, 0, st>>>((const bf16*)in, (float*)out, A, Bd, accumulate);
    else return set_error(S2S_ERR_INVALID, "transpose_last2: bad dtypes %d -> %d", in_dtype, out_dtype);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_embed_pe_fwd(const int64_t* tokens, const int32_t* ilens, const float* weight, const float* pe,
                                const float* alpha, void* y, int B, int T_in, int T_out, int d, int eos, int padding_idx,
                                const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(tokens && ilens && weight && pe && alpha && y && B > 0 && T_in > 0 && T_out > 0 && d > 0, "embed_pe_fwd: bad arguments");
    long n = (long)B * T_out * d;
    Dropout dr = make_dropout(drop);
    S2S_DISPATCH_DTYPE(dtype, TT, (embed_pe_fwd_kernel<TT><<<ew_grid(n, 1024), 256, 0, (cudaStream_t)stream>>>(
        tokens, ilens, weight, pe, alpha, (TT*)y, B, T_in, T_out, d, eos, padding_idx, dr)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_embed_pe_bwd(const void* dy, const int64_t* tokens, const int32_t* ilens, const float* pe, float* dweight,
                                float* dalpha, int B, int T_in, int T_out, int d, int eos, int padding_idx,
                                const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(dy && tokens && ilens && pe && B > 0 && T_in > 0 && T_out > 0 && d > 0, "embed_pe_bwd: bad arguments");
    long n = (long)B * T_out * d;
    Dropout dr = make_dropout(drop);
    S2S_DISPATCH_DTYPE(dtype, TT, (embed_pe_bwd_kernel<TT><<<ew_grid(n, 2048), 256, 0, (cudaStream_t)stream>>>(
        (const TT*)dy, tokens, ilens, pe, dweight, dalpha, B, T_in, T_out, d, eos, padding_idx, dr)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}


// =============================================================================================
// Dataset statistics (bin/compute_statistics.py:128-132): per-feature sum, sum of squares and row count over the valid frames of
// a zero-padded batch, accumulated in float64.  A CTA owns a contiguous slab of rows; thread t < RPI * D always works on column
// t % D (RPI = 256 / D rows per iteration), so the loads of one iteration are contiguous and the accumulators stay in registers.
// HBM-bound: 4 * D bytes per valid frame, read once.
// =============================================================================================
namespace s2s {
// V = 4: D % 4 == 0 and 16-byte aligned rows, every thread owns four consecutive columns (one float4 per row); V = 1: any D.
template <int V>
__global__ void __launch_bounds__(256) feat_stats_kernel(const float* __restrict__ x, const int* __restrict__ lens, double* __restrict__ acc,
                                                         long rows, int T, int D, int rows_per_cta) {
    __shared__ double ssum[256 * V], ssq[256 * V];
    __shared__ unsigned long long scount;
    const int G = D / V;                                  // column groups of the whole row
    const int g0 = blockIdx.y * 256;                      // column-group tile (G > 256)
    const int Gt = min(256, G - g0);
    const int rpi = 256 / Gt;
    const int cg = threadIdx.x % Gt, rl = threadIdx.x / Gt;
    const bool active = rl < rpi;
    if (threadIdx.x == 0) scount = 0ull;
    __syncthreads();
    double s1[V], s2[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { s1[k] = 0.0; s2[k] = 0.0; }
    unsigned long long cnt = 0;
    const long r0 = (long)blockIdx.x * rows_per_cta;
    const long r1 = min(rows, r0 + rows_per_cta);
    if (active) {
        // (utterance, frame) of the row advance incrementally: one division per thread, not one per row
        long r = r0 + rl;
        int b = (int)(r / T), t = (int)(r - (long)b * T), cur_b = -1, len = T;
        for (; r < r1; r += rpi, t += rpi) {
            while (t >= T) { t -= T; ++b; }
            if (lens && b != cur_b) { len = lens[b]; cur_b = b; }
            if (t >= len) continue;
            float v[V];
            if constexpr (V == 4) {
                const float4 q = *reinterpret_cast<const float4*>(x + r * D + (long)(g0 + cg) * 4);
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
                v[0] = x[r * D + g0 + cg];
            }
#pragma unroll
            for (int k = 0; k < V; ++k) { s1[k] += (double)v[k]; s2[k] += (double)v[k] * (double)v[k]; }
            ++cnt;
        }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) { ssum[k * 256 + threadIdx.x] = s1[k]; ssq[k * 256 + threadIdx.x] = s2[k]; }
    if (active && cg == 0 && blockIdx.y == 0 && cnt) atomicAdd(&scount, cnt);
    __syncthreads();
    for (int i = threadIdx.x; i < Gt * V; i += 256) {
        const int g = i / V, k = i % V;
        double a = 0.0, b = 0.0;
        for (int j = 0; j < rpi; ++j) { a += ssum[k * 256 + j * Gt + g]; b += ssq[k * 256 + j * Gt + g]; }
        const int c = (g0 + g) * V + k;
        atomicAdd(acc + c, a);
        atomicAdd(acc + D + c, b);
    }
    if (threadIdx.x == 0 && blockIdx.y == 0 && scount) atomicAdd(acc + 2 * D, (double)scount);
}
}  // namespace s2s

extern "C" int s2s_feat_stats(const float* feats, const int* lens, double* acc, int B, int T, int D, void* stream) {
    using namespace s2s;
    S2S_REQUIRE(feats && acc && B >= 0 && T >= 0 && D > 0, "feat_stats: bad arguments");
    const long rows = (long)B * T;
    if (rows == 0) return S2S_OK;
    const long want = (long)num_sms() * 8;
    long per = ceil_div_l(rows, want);
    if (per < 64) per = 64;
    const unsigned gx = (unsigned)ceil_div_l(rows, per);
    if (D % 4 == 0 && (reinterpret_cast<uintptr_t>(feats) & 15) == 0)
        feat_stats_kernel<4><<<dim3(gx, (unsigned)ceil_div_l(D / 4, 256)), 256, 0, (cudaStream_t)stream>>>(feats, lens, acc, rows, T, D, (int)per);
    else
        feat_stats_kernel<1><<<dim3(gx, (unsigned)ceil_div_l(D, 256)), 256, 0, (cudaStream_t)stream>>>(feats, lens, acc, rows, T, D, (int)per);
    S2S_LAUNCH_OK();
    return S2S_OK;
}


// =============================================================================================
// Generic square-kernel / stride im2col and col2im over channels-last maps (the second and third convolutions of
// Conv2dSubsampling2 / 6 / 8, subsampling.py:108-279: (k, s) = (3, 1), (5, 3), (3, 2)); the (3, 2) hot path of Conv2dSubsampling
// keeps its specialised kernels above.  col[(b, t2, f2), tap, c] = y[b, s t2 + kt, s f2 + kf, c], T2 = (T1 - k) / s + 1.
// =============================================================================================
namespace s2s {
template <typename T, int VEC>
__global__ void __launch_bounds__(256) im2col2d_kernel(const T* __restrict__ y, T* __restrict__ col, int B, int T1, int F1, int C, int T2, int F2,
                                                       int k, int s) {
    const int cv = C / VEC, kk = k * k;
    const long total = (long)B * T2 * F2 * kk * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * VEC;
        const long q = i / cv;
        const int tap = (int)(q % kk);
        const long m = q / kk;
        const int f2 = (int)(m % F2);
        const long r = m / F2;
        const int t2 = (int)(r % T2);
        const long b = r / T2;
        const int kt = tap / k, kf = tap - kt * k;
        const T* src = y + (((b * T1 + (long)s * t2 + kt) * F1) + (long)s * f2 + kf) * C + c;
        T* dst = col + (m * kk + tap) * C + c;
        if constexpr (VEC == 8) {
            float v[8];
            Vec8<T>::load(src, v);
            Vec8<T>::store(dst, v);
        } else {
            *dst = *src;
        }
    }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) col2im2d_kernel(const T* __restrict__ dcol, T* __restrict__ dy, int B, int T1, int F1, int C, int T2, int F2,
                                                       int k, int s, const T* __restrict__ gate) {
    const int cv = C / VEC, kk = k * k;
    const long total = (long)B * T1 * F1 * cv;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv) * VEC;
        const long q = i / cv;
        const int f1 = (int)(q % F1);
        const long r = q / F1;
        const int t1 = (int)(r % T1);
        const long b = r / T1;
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        for (int kt = 0; kt < k; ++kt) {
            const int tt = t1 - kt;
            if (tt < 0 || tt % s) continue;
            const int t2 = tt / s;
            if (t2 >= T2) continue;
            for (int kf = 0; kf < k; ++kf) {
                const int ff = f1 - kf;
                if (ff < 0 || ff % s) continue;
                const int f2 = ff / s;
                if (f2 >= F2) continue;
                const long m = (b * T2 + t2) * F2 + f2;
                const T* src = dcol + (m * kk + kt * k + kf) * C + c;
                if constexpr (VEC == 8) {
                    float v[8];
                    Vec8<T>::load(src, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] += v[j];
                } else {
                    acc[0] += to_f<T>(*src);
                }
            }
        }
        T* dst = dy + q * C + c;
        if (gate) {                                   // ReLU' of the layer that produced the map
            if constexpr (VEC == 8) {
                float g[8];
                Vec8<T>::load(gate + q * C + c, g);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = g[j] > 0.f ? acc[j] : 0.f;
            } else {
                if (!(to_f<T>(gate[q * C + c]) > 0.f)) acc[0] = 0.f;
            }
        }
        if constexpr (VEC == 8) Vec8<T>::store(dst, acc);
        else *dst = from_f<T>(acc[0]);
    }
}
}  // namespace s2s

extern "C" int s2s_im2col2d(const void* y, void* col, int B, int T1, int F1, int C, int k, int s, int dtype, void* stream) {
    S2S_REQUIRE(y && col && B > 0 && k >= 1 && s >= 1 && T1 >= k && F1 >= k && C > 0, "im2col2d: bad arguments");
    const int T2 = (T1 - k) / s + 1, F2 = (F1 - k) / s + 1;
    const long total = (long)B * T2 * F2 * k * k * C;
    const bool ok = (C % 8 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)col % 16 == 0);
    S2S_DISPATCH_DTYPE(dtype, TT, {
        if (ok) im2col2d_kernel<TT, 8><<<ew_grid(total / 8, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)y, (TT*)col, B, T1, F1, C, T2, F2, k, s);
        else im2col2d_kernel<TT, 1><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)y, (TT*)col, B, T1, F1, C, T2, F2, k, s);
    });
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_col2im2d(const void* dcol, const void* gate, void* dy, int B, int T1, int F1, int C, int k, int s, int dtype, void* stream) {
    S2S_REQUIRE(dcol && dy && B > 0 && k >= 1 && s >= 1 && T1 >= k && F1 >= k && C > 0, "col2im2d: bad arguments");
    const int T2 = (T1 - k) / s + 1, F2 = (F1 - k) / s + 1;
    const long total = (long)B * T1 * F1 * C;
    const bool ok = (C % 8 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)dcol % 16 == 0) && ((uintptr_t)gate % 16 == 0);
    S2S_DISPATCH_DTYPE(dtype, TT, {
        if (ok) col2im2d_kernel<TT, 8><<<ew_grid(total / 8, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)dcol, (TT*)dy, B, T1, F1, C, T2, F2, k, s, (const TT*)gate);
        else col2im2d_kernel<TT, 1><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const TT*)dcol, (TT*)dy, B, T1, F1, C, T2, F2, k, s, (const TT*)gate);
    });
    S2S_LAUNCH_OK();
    return S2S_OK;
}


// =============================================================================================
// Weight gradient of the first convolution (Conv2d(1 -> C, 3, 2)) as a tensor-core product (bf16 engine):
//   dW[c, tap] = sum_p dy1[p, c] * x[p @ tap],  dbias[c] = sum_p dy1[p, c]      p = (b, t1, f1): 318 k positions at the C2 shape
// is dy1^T (C x P) times a (P x 16) patch matrix of the INPUT (nine taps, a column of ones for the bias, zero padding): 10 MB instead
// of the CUDA-core kernel's 80 FMAs per position and thread (conv1_bwd8_kernel: 189 us at the C2 shape, issue-bound; the product
// streams dy1 once through s2s_gemm).  conv1_xcol_kernel builds the patch matrix, conv1_dw_scatter_kernel adds the (C, 16) result
// into the parameter gradients.
// =============================================================================================
namespace s2s {
template <typename T>
__global__ void __launch_bounds__(256) conv1_xcol_kernel(const float* __restrict__ x, T* __restrict__ xcol, long P, int Tn, int F, int T1, int F1) {
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
        const int f1 = (int)(p % F1);
        const long q = p / F1;
        const int t1 = (int)(q % T1);
        const long b = q / T1;
        const float* xp = x + (b * Tn + 2 * t1) * F + 2 * f1;
        float v[16];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt)
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) v[kt * 3 + kf] = xp[kt * F + kf];
        v[9] = 1.f;
#pragma unroll
        for (int k = 10; k < 16; ++k) v[k] = 0.f;
        T* dst = xcol + p * 16;
        Vec8<T>::store(dst, reinterpret_cast<float(&)[8]>(v[0]));
        Vec8<T>::store(dst + 8, reinterpret_cast<float(&)[8]>(v[8]));
    }
}
// w16[c] = [w[c, 0..8] | bias[c] | 0 x 6]: the first convolution's weights as the (C, 16) operand of patches x w16^T (+ bias through the
// ones column of the patch matrix)
template <typename T>
__global__ void conv1_pack_w_kernel(const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ w16, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * 16) return;
    const int c = i >> 4, k = i & 15;
    w16[i] = from_f<T>(k < 9 ? w[c * 9 + k] : (k == 9 ? bias[c] : 0.f));
}
__global__ void conv1_dw_scatter_kernel(const float* __restrict__ g16, float* __restrict__ dw, float* __restrict__ dbias, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * 10) return;
    const int c = i / 10, k = i - c * 10;
    const float v = g16[c * 16 + k];
    if (k < 9) dw[c * 9 + k] += v;
    else dbias[c] += v;
}
}  // namespace s2s

extern "C" int s2s_conv1_xcol(const float* x, void* xcol, int B, int T, int F, int dtype, void* stream) {
    S2S_REQUIRE(x && xcol && B > 0 && T >= 3 && F >= 3, "conv1_xcol: bad arguments");
    const int T1 = (T - 1) / 2, F1 = (F - 1) / 2;
    const long P = (long)B * T1 * F1;
    S2S_DISPATCH_DTYPE(dtype, TT, (conv1_xcol_kernel<TT><<<ew_grid(P, 256), 256, 0, (cudaStream_t)stream>>>(x, (TT*)xcol, P, T, F, T1, F1)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_conv1_pack_w(const float* w, const float* bias, void* w16, int C, int dtype, void* stream) {
    S2S_REQUIRE(w && bias && w16 && C > 0, "conv1_pack_w: bad arguments");
    S2S_DISPATCH_DTYPE(dtype, TT, (conv1_pack_w_kernel<TT><<<(unsigned)ceil_div_l((long)C * 16, 256), 256, 0, (cudaStream_t)stream>>>(w, bias, (TT*)w16, C)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_conv1_dw_scatter(const float* g16, float* dw, float* dbias, int C, void* stream) {
    S2S_REQUIRE(g16 && dw && dbias && C > 0, "conv1_dw_scatter: bad arguments");
    conv1_dw_scatter_kernel<<<(unsigned)ceil_div_l((long)C * 10, 256), 256, 0, (cudaStream_t)stream>>>(g16, dw, dbias, C);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
