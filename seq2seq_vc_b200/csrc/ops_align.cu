// AAS-VC alignment block (reference: seq2seq_vc/modules/alignments.py:28-60 AlignmentModule distance + log-softmax,
// losses/forward_sum_loss.py:26-76 ForwardSumLoss (= per-utterance CTC over a constant-blank lattice),
// modules/length_regulator.py:111-154 GaussianUpsampling weights, losses/duration_predictor_loss.py:29-50).
//
//  * pairwise L2 distance is computed by DIRECT differences in fp32 on the CUDA cores (no ||f||^2+||t||^2-2ft
//    cancellation), tiled 64x64 per CTA, and never materialises the reference's (B,T_feats,T_text,C) tensor:
//    HBM traffic is read (T_feats+T_text)*C, write T_feats*T_text per utterance;
//  * forward-sum runs the alpha and beta recursions of one utterance in one CTA (one thread per lattice state,
//    one __syncthreads per frame, log-domain fp32) and emits the gradient that torch's ctc_loss backward hands the
//    reference: exp(lp) - exp(alpha+beta-lp+nll)  (aten LossCTC.cpp "eq. 16"; the exp(lp) term does not cancel
//    because the reference's rows are un-normalised - it is part of the reference's training signal).
#include "common.cuh"

namespace s2s {

// ---------------------------------------------------------------------------------------------
// dist[b,t,s] = || f[b,t,:] - x[b,s,:] ||_2      f: (B,Tf,C)  x: (B,Tt,C)  dist: (B,Tf,Tt) f32
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pairdist_kernel(const T* __restrict__ f, const T* __restrict__ x, float* __restrict__ dist,
                                                       int Tf, int Tt, int C) {
    constexpr int BM = 64, BN = 64, BK = 32;
    __shared__ float Fs[BK][BM + 1];
    __shared__ float Xs[BK][BN + 1];
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * BM, s0 = blockIdx.x * BN;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const T* fb = f + (long)b * Tf * C;
    const T* xb = x + (long)b * Tt * C;
    // thread (tx, ty) owns the 4 x 4 block of frames ty * 4 + i and tokens tx * 4 + j.  Measured and rejected (C3 shape, 1.42 ms
    // for this form): an interleaved mapping (frames ty + 16 i, tokens tx + 16 j) that removes the 2-way bank conflict on the
    // token tile -- 1.51 ms; the same with packed FADD2 / FFMA2 accumulators -- 1.58 ms (124 registers, and the packed ops do not
    // raise the fp32 pipe's throughput).
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lk = tid & 31, lr = tid >> 5;   // loader: 8 rows x 32 k per pass
    for (int k0 = 0; k0 < C; k0 += BK) {
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int r = lr + p * 8;
            const int k = k0 + lk;
            const int tr = t0 + r, sr = s0 + r;
            Fs[lk][r] = (tr < Tf && k < C) ? to_f<T>(fb[(long)tr * C + k]) : 0.f;
            Xs[lk][r] = (sr < Tt && k < C) ? to_f<T>(xb[(long)sr * C + k]) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < BK; ++k) {
            float a[4], c[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Fs[k][ty * 4 + i]; c[i] = Xs[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float d = a[i] - c[j]; acc[i][j] = fmaf(d, d, acc[i][j]); }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty * 4 + i;
        if (t >= Tf) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int s = s0 + tx * 4 + j;
            if (s < Tt) dist[((long)b * Tf + t) * Tt + s] = sqrtf(acc[i][j]);
        }
    }
}

// in place: logp[b,t,s] = -dist - lse_row  (s < text_len[b]), -inf otherwise; lse[b,t] = log sum_s exp(-dist).  warp per row
__global__ void __launch_bounds__(256) neg_logsoftmax_kernel(float* __restrict__ logp, float* __restrict__ lse,
                                                             const int32_t* __restrict__ text_lens, long rows, int Tf, int Tt) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = (int)(row / Tf);
    int n = text_lens[b];
    n = n < 0 ? 0 : (n > Tt ? Tt : n);
    float* p = logp + row * Tt;
    float mx = -INFINITY;
    for (int s = lane; s < n; s += 32) mx = fmaxf(mx, -p[s]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < n; s += 32) sum += expf(-p[s] - mx);
    sum = warp_sum(sum);
    const float l = (n > 0) ? mx + logf(sum) : 0.f;
    for (int s = lane; s < Tt; s += 32) p[s] = (s < n) ? -p[s] - l : -INFINITY;
    if (lane == 0) lse[row] = l;
}

// out[r] = sum_c x[r,c]^2 (fp32), warp per row
template <typename T>
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const T* __restrict__ x, float* __restrict__ out, long rows, int C) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* p = x + row * C;
    float acc = 0.f;
    if ((C & 7) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        for (int c = lane * 8; c < C; c += 256) {
            float v[8];
            Vec8<T>::load(p + c, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fmaf(v[k], v[k], acc);
        }
    } else {
        for (int c = lane; c < C; c += 32) { const float v = to_f<T>(p[c]); acc = fmaf(v, v, acc); }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
}

// The same log-probabilities from the tensor-core product: logp holds f . x (B,Tf,Tt) on entry; dist = sqrt(max(|f|^2 + |x|^2 - 2 f.x, 0)),
// then the masked log-softmax of -dist in place.  Used by the bf16 engine only: its operands are bf16 values, so the products are
// exact in the fp32 accumulator and the expansion's cancellation error (~1e-7 |f|^2 in dist^2) is far below the operands' rounding; the
// float32 parity path keeps the direct-difference kernel above.  warp per row
__global__ void __launch_bounds__(256) dot_to_logp_kernel(float* __restrict__ logp, const float* __restrict__ nf, const float* __restrict__ nx,
                                                          float* __restrict__ lse, const int32_t* __restrict__ text_lens, long rows, int Tf,
                                                          int Tt) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = (int)(row / Tf);
    int n = text_lens[b];
    n = n < 0 ? 0 : (n > Tt ? Tt : n);
    float* p = logp + row * Tt;
    const float* nxb = nx + (long)b * Tt;
    const float a = nf[row];
    float mx = -INFINITY;
    for (int s = lane; s < n; s += 32) {
        const float d = sqrtf(fmaxf(a + nxb[s] - 2.f * p[s], 0.f));
        p[s] = d;
        mx = fmaxf(mx, -d);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < n; s += 32) sum += expf(-p[s] - mx);
    sum = warp_sum(sum);
    const float l = (n > 0) ? mx + logf(sum) : 0.f;
    for (int s = lane; s < Tt; s += 32) p[s] = (s < n) ? -p[s] - l : -INFINITY;
    if (lane == 0) lse[row] = l;
}

// backward of (-dist -> log_softmax): W[b,t,s] = d_dist / dist with d_dist = -(dlogp - exp(logp) * sum_s dlogp);
// rowsum[b,t] = sum_s W.  W: (B,Tf,ldW) activation dtype, columns >= text_len are zero.  warp per row
template <typename T>
__global__ void __launch_bounds__(256) align_bwd_kernel(const float* __restrict__ dlogp, const float* __restrict__ logp,
                                                        const float* __restrict__ lse, const int32_t* __restrict__ text_lens,
                                                        T* __restrict__ W, float* __restrict__ rowsum, long rows, int Tf, int Tt,
                                                        long ldW) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int b = (int)(row / Tf);
    int n = text_lens[b];
    n = n < 0 ? 0 : (n > Tt ? Tt : n);
    const float* g = dlogp + row * Tt;
    const float* p = logp + row * Tt;
    const float l = lse[row];
    float G = 0.f;
    for (int s = lane; s < n; s += 32) G += g[s];
    G = warp_sum(G);
    float rs = 0.f;
    T* w = W + row * ldW;
    for (int s = lane; s < (int)ldW; s += 32) {
        float o = 0.f;
        if (s < n) {
            const float lp = p[s];
            const float dscore = g[s] - expf(lp) * G;
            const float dist = -(lp + l);
            o = dist > 0.f ? -dscore / dist : 0.f;
            o = to_f<T>(from_f<T>(o));          // rowsum must see the value the GEMM will see
            rs += o;
        }
        w[s] = from_f<T>(o);
    }
    rs = warp_sum(rs);
    if (lane == 0) rowsum[row] = rs;
}

// colsum[b,s] = sum_t W[b,t,s]    grid (ceil(Tt/32), B), block (32, 8)
template <typename T>
__global__ void __launch_bounds__(256) batched_colsum_kernel(const T* __restrict__ W, float* __restrict__ out, int Tf, int Tt, long ldW) {
    __shared__ float red[8][32];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int s = blockIdx.x * 32 + tx;
    const int b = blockIdx.y;
    float acc = 0.f;
    if (s < Tt)
        for (int t = ty; t < Tf; t += 8) acc += to_f<T>(W[((long)b * Tf + t) * ldW + s]);
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && s < Tt) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red[q][tx];
        out[(long)b * Tt + s] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// forward-sum (CTC) loss + gradient, one CTA per utterance, thread s <-> lattice state s of [blank,1,blank,...,N,blank]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lse3f(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    if (m == -INFINITY) return -INFINITY;
    return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__global__ void __launch_bounds__(1024) forward_sum_kernel(const float* __restrict__ logp, const float* __restrict__ prior,
                                                           const int32_t* __restrict__ text_lens,
                                                           const int32_t* __restrict__ feats_lens, int B, int Tf, int Tt,
                                                           float blank_logp, float* __restrict__ alpha_ws, float* __restrict__ loss,
                                                           float* __restrict__ dlogp, float gscale) {
    extern __shared__ float sm[];                 // 2 x (Smax + 2), two leading -inf guards per buffer
    __shared__ float s_ll;
    const int b = blockIdx.x;
    const int s = threadIdx.x;
    int N = text_lens[b], T = feats_lens[b];
    N = N < 0 ? 0 : (N > Tt ? Tt : N);
    T = T < 0 ? 0 : (T > Tf ? Tf : T);
    const int S = 2 * N + 1;
    const int Smax = 2 * Tt + 1;
    const int stride = Smax + 4;
    float* buf0 = sm + 2;
    float* buf1 = sm + stride + 2;
    const float* lp_b = logp + (long)b * Tf * Tt;
    const float* pr_b = prior + (long)b * Tf * Tt;
    float* al_b = alpha_ws + (long)b * Tf * Tt;
    float* g_b = dlogp ? dlogp + (long)b * Tf * Tt : nullptr;
    const bool is_label = (s & 1) && s < S;
    const int k = s >> 1;
    const bool active = s < S;

    if (N == 0 || T == 0) {                        // nothing to align: zero gradient block, no loss
        if (g_b)
            for (long i = s; i < (long)Tf * Tt; i += blockDim.x) g_b[i] = 0.f;
        return;
    }
    if (s < 2) { buf0[-1 - s] = -INFINITY; buf1[-1 - s] = -INFINITY; }
    // ---- alpha
    float e = is_label ? lp_b[k] + pr_b[k] : blank_logp;
    float a = (s < 2 && active) ? e : -INFINITY;
    if (s <= Smax + 1) buf0[s] = active ? a : -INFINITY;
    if (is_label) al_b[k] = a;
    __syncthreads();
    float* cur = buf0;
    float* nxt = buf1;
    e = (is_label && T > 1) ? lp_b[(long)Tt + k] + pr_b[(long)Tt + k] : blank_logp;
    for (int t = 1; t < T; ++t) {
        const float e_t = e;
        if (is_label && t + 1 < T) e = lp_b[(long)(t + 1) * Tt + k] + pr_b[(long)(t + 1) * Tt + k];   // prefetch next frame
        if (active) {
            const float a0 = cur[s], a1 = cur[s - 1];
            const float a2 = (is_label && s >= 3) ? cur[s - 2] : -INFINITY;
            a = lse3f(a0, a1, a2) + e_t;
            nxt[s] = a;
            if (is_label) al_b[(long)t * Tt + k] = a;
        }
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (s == 0) {
        const float x = cur[S - 1], y = (S >= 2) ? cur[S - 2] : -INFINITY;
        s_ll = lse3f(x, y, -INFINITY);
    }
    __syncthreads();
    const float ll = s_ll;
    const bool feasible = ll > -INFINITY && ll < INFINITY;
    if (s == 0 && feasible) atomicAdd(loss, -ll / ((float)N * (float)B));
    if (!g_b) return;
    // ---- zero the part of the gradient block this utterance does not cover
    for (long i = s; i < (long)Tf * Tt; i += blockDim.x) {
        const int tt = (int)(i / Tt), kk = (int)(i - (long)tt * Tt);
        if (!feasible || tt >= T || kk >= N) g_b[i] = 0.f;
    }
    if (!feasible) return;
    // ---- beta (guards: two trailing -inf after state S-1)
    __syncthreads();
    const float sc = gscale / ((float)N * (float)B);
    const float nll = -ll;
    e = is_label ? lp_b[(long)(T - 1) * Tt + k] + pr_b[(long)(T - 1) * Tt + k] : blank_logp;
    float bt = (active && s >= S - 2) ? e : -INFINITY;
    if (s <= Smax + 1) cur[s] = active ? bt : -INFINITY;     // states >= S read as -inf
    if (is_label) {
        const float al = al_b[(long)(T - 1) * Tt + k];
        g_b[(long)(T - 1) * Tt + k] = sc * (expf(e) - expf(al + bt - e + nll));
    }
    __syncthreads();
    for (int t = T - 2; t >= 0; --t) {
        e = is_label ? lp_b[(long)t * Tt + k] + pr_b[(long)t * Tt + k] : blank_logp;
        if (active) {
            const float b0 = cur[s];
            const float b1 = (s + 1 < S) ? cur[s + 1] : -INFINITY;
            const float b2 = (is_label && s + 2 < S) ? cur[s + 2] : -INFINITY;
            bt = lse3f(b0, b1, b2) + e;
            nxt[s] = bt;
            if (is_label) {
                const float al = al_b[(long)t * Tt + k];
                g_b[(long)t * Tt + k] = sc * (expf(e) - expf(al + bt - e + nll));
            }
        }
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
}

// Same recursions with the alpha and the beta sweep running CONCURRENTLY in one CTA (threads [0, SH) walk frames 0 .. T-1,
// threads [SH, 2 SH) walk frames T-1 .. 0, one __syncthreads per frame serves both), each sweep fetching its next frame's
// emissions one step ahead: the serial chain is T steps instead of 2 T.  Beta is parked in the gradient block and a final
// fully parallel pass turns (alpha, beta) into torch's ctc_loss gradient in place.  Used when 2 SH <= 1024 and a gradient is wanted.
__global__ void __launch_bounds__(1024) forward_sum_par_kernel(const float* __restrict__ logp, const float* __restrict__ prior,
                                                               const int32_t* __restrict__ text_lens,
                                                               const int32_t* __restrict__ feats_lens, int B, int Tf, int Tt, int SH,
                                                               float blank_logp, float* __restrict__ alpha_ws, float* __restrict__ loss,
                                                               float* __restrict__ dlogp, float gscale) {
    extern __shared__ float sm[];                 // 4 x (Smax + 4): alpha cur / next, beta cur / next, two -inf guards in front of each
    __shared__ float s_ll;
    const int b = blockIdx.x;
    const bool beta_half = (int)threadIdx.x >= SH;
    const int s = beta_half ? (int)threadIdx.x - SH : (int)threadIdx.x;
    int N = text_lens[b], T = feats_lens[b];
    N = N < 0 ? 0 : (N > Tt ? Tt : N);
    T = T < 0 ? 0 : (T > Tf ? Tf : T);
    const int S = 2 * N + 1;
    const int Smax = 2 * Tt + 1;
    const int stride = Smax + 4;
    float* cur = sm + (beta_half ? 2 * stride : 0) + 2;
    float* nxt = cur + stride;
    const float* lp_b = logp + (long)b * Tf * Tt;
    const float* pr_b = prior + (long)b * Tf * Tt;
    float* al_b = alpha_ws + (long)b * Tf * Tt;
    float* g_b = dlogp + (long)b * Tf * Tt;
    const bool is_label = (s & 1) && s < S;
    const int k = s >> 1;
    const bool active = s < S;
    const long total = (long)Tf * Tt;
    if (N == 0 || T == 0) {
        for (long i = threadIdx.x; i < total; i += blockDim.x) g_b[i] = 0.f;
        return;
    }
    if (s < 2) { cur[-1 - s] = -INFINITY; nxt[-1 - s] = -INFINITY; }
    // ---- first frame of each sweep
    const int t0 = beta_half ? T - 1 : 0;
    float e = is_label ? lp_b[(long)t0 * Tt + k] + pr_b[(long)t0 * Tt + k] : blank_logp;
    float a;
    if (!beta_half) a = (s < 2 && active) ? e : -INFINITY;
    else a = (active && s >= S - 2) ? e : -INFINITY;
    if (s <= Smax + 1) cur[s] = active ? a : -INFINITY;
    if (is_label) (beta_half ? g_b : al_b)[(long)t0 * Tt + k] = a;
    __syncthreads();
    int tn = beta_half ? T - 2 : 1;               // frame of the next step, its emission is fetched one step ahead
    e = (is_label && T > 1) ? lp_b[(long)tn * Tt + k] + pr_b[(long)tn * Tt + k] : blank_logp;
    for (int j = 1; j < T; ++j) {
        const int t = tn;
        const float e_t = e;
        tn = beta_half ? t - 1 : t + 1;
        if (is_label && j + 1 < T) e = lp_b[(long)tn * Tt + k] + pr_b[(long)tn * Tt + k];
        if (active) {
            float x0, x1, x2;
            if (!beta_half) {
                x0 = cur[s]; x1 = cur[s - 1];
                x2 = (is_label && s >= 3) ? cur[s - 2] : -INFINITY;
            } else {
                x0 = cur[s];
                x1 = (s + 1 < S) ? cur[s + 1] : -INFINITY;
                x2 = (is_label && s + 2 < S) ? cur[s + 2] : -INFINITY;
            }
            a = lse3f(x0, x1, x2) + e_t;
            nxt[s] = a;
            if (is_label) (beta_half ? g_b : al_b)[(long)t * Tt + k] = a;
        }
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (threadIdx.x == 0) {                        // alpha half, after the last frame
        const float x = cur[S - 1], y = (S >= 2) ? cur[S - 2] : -INFINITY;
        s_ll = lse3f(x, y, -INFINITY);
    }
    __syncthreads();                               // also orders the parked alpha / beta values before the pass below
    const float ll = s_ll;
    const bool feasible = ll > -INFINITY && ll < INFINITY;
    if (threadIdx.x == 0 && feasible) atomicAdd(loss, -ll / ((float)N * (float)B));
    const float sc = gscale / ((float)N * (float)B);
    const float nll = -ll;
    for (long i = threadIdx.x; i < total; i += blockDim.x) {
        const int tt = (int)(i / Tt), kk = (int)(i - (long)tt * Tt);
        float g = 0.f;
        if (feasible && tt < T && kk < N) {
            const float ee = lp_b[i] + pr_b[i];
            g = sc * (expf(ee) - expf(al_b[i] + g_b[i] - ee + nll));
        }
        g_b[i] = g;
    }
}

// ---------------------------------------------------------------------------------------------
// Gaussian upsampling weights: P[b,t,s] = softmax_s( -delta * (t_eff - c_s)^2 ), c = cumsum(ds) - ds/2,
// t_eff = t for t < feats_len[b] else 0 (reference quirk), text padding masked.   P: (B,Tf,ldP)
// grid (ceil(Tf/8), B), block 256 (8 warps = 8 rows)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gauss_weights_kernel(const float* __restrict__ ds, const int32_t* __restrict__ feats_lens,
                                                            const int32_t* __restrict__ text_lens, T* __restrict__ P, int Tf, int Tt,
                                                            long ldP, float delta) {
    extern __shared__ float c[];
    const int b = blockIdx.y;
    int n = text_lens[b];
    n = n < 0 ? 0 : (n > Tt ? Tt : n);
    if (threadIdx.x == 0) {
        float run = 0.f;
        for (int s = 0; s < Tt; ++s) { const float d = ds[(long)b * Tt + s]; run += d; c[s] = run - d / 2.f; }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= Tf) return;
    const float te = (t < feats_lens[b]) ? (float)t : 0.f;
    const float nd = -1.f * delta;
    float mx = -INFINITY;
    for (int s = lane; s < n; s += 32) { const float d = te - c[s]; mx = fmaxf(mx, nd * (d * d)); }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < n; s += 32) { const float d = te - c[s]; sum += expf(nd * (d * d) - mx); }
    sum = warp_sum(sum);
    const float inv = n > 0 ? 1.f / sum : 0.f;
    T* p = P + ((long)b * Tf + t) * ldP;
    for (int s = lane; s < (int)ldP; s += 32) {
        float o = 0.f;
        if (s < n) { const float d = te - c[s]; o = expf(nd * (d * d) - mx) * inv; }
        p[s] = from_f<T>(o);
    }
}

// ---------------------------------------------------------------------------------------------
// duration loss: d_out = min(pre * mask, 10); loss = mean_valid (d_out - log(ds + offset))^2; d_pre. one CTA
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) duration_loss_kernel(const T* __restrict__ pre, const float* __restrict__ ds,
                                                            const int32_t* __restrict__ text_lens, int B, int Tt, float offset,
                                                            float clamp_max, float gscale, const float* __restrict__ g_douts,
                                                            float* __restrict__ d_outs, float* __restrict__ loss,
                                                            T* __restrict__ d_pre) {
    __shared__ float red[32];
    __shared__ float s_n;
    float cnt = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) { int n = text_lens[b]; cnt += (float)(n < 0 ? 0 : (n > Tt ? Tt : n)); }
    cnt = block_sum(cnt, red);
    if (threadIdx.x == 0) s_n = cnt;
    __syncthreads();
    const float nv = s_n;
    float acc = 0.f;
    for (int i = threadIdx.x; i < B * Tt; i += blockDim.x) {
        const int b = i / Tt, s = i - b * Tt;
        const bool valid = s < text_lens[b];
        const float x = valid ? to_f<T>(pre[i]) : 0.f;
        const float d = fminf(x, clamp_max);
        if (d_outs) d_outs[i] = d;
        float g = 0.f;
        if (valid) {
            const float diff = d - logf(ds[i] + offset);
            acc += diff * diff;
            if (x <= clamp_max && nv > 0.f) g = g_douts ? g_douts[i] : gscale * 2.f * diff / nv;
        }
        if (d_pre) d_pre[i] = from_f<T>(g);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0 && loss) *loss = nv > 0.f ? acc / nv : 0.f;
}

// inference durations: d = min(max(rint(exp(pre) - offset), 0), clamp_max)   (duration_predictor.py:92-96, aas_vc.py:389)
template <typename T>
__global__ void duration_infer_kernel(const T* __restrict__ pre, float* __restrict__ d, int n, float offset, float clamp_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = rintf(expf(to_f<T>(pre[i])) - offset);
    v = fmaxf(v, 0.f);
    d[i] = fminf(v, clamp_max);
}

}  // namespace s2s

using namespace s2s;

extern "C" int s2s_align_logp_fwd(const void* feats, const void* text, const int32_t* text_lens, float* logp, float* lse, int B,
                                  int T_feats, int T_text, int C, int dtype, void* stream) {
    S2S_REQUIRE(feats && text && text_lens && logp && lse && B > 0 && T_feats > 0 && T_text > 0 && C > 0 && B <= 65535,
                "align_logp_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div_l(T_text, 64), (unsigned)ceil_div_l(T_feats, 64), (unsigned)B);
    S2S_DISPATCH_DTYPE(dtype, T, (pairdist_kernel<T><<<grid, 256, 0, st>>>((const T*)feats, (const T*)text, logp, T_feats, T_text, C)));
    S2S_LAUNCH_OK();
    long rows = (long)B * T_feats;
    neg_logsoftmax_kernel<<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(logp, lse, text_lens, rows, T_feats, T_text);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_row_sqnorm(const void* x, float* out, int64_t rows, int C, int dtype, void* stream) {
    S2S_REQUIRE(x && out && rows > 0 && C > 0, "row_sqnorm: bad arguments");
    S2S_DISPATCH_DTYPE(dtype, T, (row_sqnorm_kernel<T><<<(unsigned)ceil_div_l(rows, 8), 256, 0, (cudaStream_t)stream>>>((const T*)x, out, rows, C)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_align_logp_from_dot(float* logp, const float* feats_sqnorm, const float* text_sqnorm, const int32_t* text_lens, float* lse,
                                       int B, int T_feats, int T_text, void* stream) {
    S2S_REQUIRE(logp && feats_sqnorm && text_sqnorm && text_lens && lse && B > 0 && T_feats > 0 && T_text > 0, "align_logp_from_dot: bad arguments");
    const long rows = (long)B * T_feats;
    dot_to_logp_kernel<<<(unsigned)ceil_div_l(rows, 8), 256, 0, (cudaStream_t)stream>>>(logp, feats_sqnorm, text_sqnorm, lse, text_lens, rows,
                                                                                      T_feats, T_text);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_align_logp_bwd(const float* dlogp, const float* logp, const float* lse, const int32_t* text_lens, void* W,
                                  float* rowsum, float* colsum, int B, int T_feats, int T_text, int64_t ldW, int dtype,
                                  void* stream) {
    S2S_REQUIRE(dlogp && logp && lse && text_lens && W && rowsum && colsum && B > 0 && T_feats > 0 && T_text > 0 && ldW >= T_text &&
                    B <= 65535, "align_logp_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long rows = (long)B * T_feats;
    S2S_DISPATCH_DTYPE(dtype, T, (align_bwd_kernel<T><<<(unsigned)ceil_div_l(rows, 8), 256, 0, st>>>(
        dlogp, logp, lse, text_lens, (T*)W, rowsum, rows, T_feats, T_text, ldW)));
    S2S_LAUNCH_OK();
    dim3 grid((unsigned)ceil_div_l(T_text, 32), (unsigned)B), block(32, 8);
    S2S_DISPATCH_DTYPE(dtype, T, (batched_colsum_kernel<T><<<grid, block, 0, st>>>((const T*)W, colsum, T_feats, T_text, ldW)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_forward_sum(const float* logp, const float* prior, const int32_t* text_lens, const int32_t* feats_lens, int B,
                               int T_feats, int T_text, float blank_logp, float* alpha_ws, float* loss, float* dlogp,
                               float grad_scale, void* stream) {
    S2S_REQUIRE(logp && prior && text_lens && feats_lens && alpha_ws && loss && B > 0 && T_feats > 0 && T_text > 0,
                "forward_sum: bad arguments");
    S2S_REQUIRE(2 * T_text + 3 <= 1024, "forward_sum: T_text %d exceeds the 510 tokens-per-CTA limit", T_text);
    cudaStream_t st = (cudaStream_t)stream;
    int threads = ((2 * T_text + 1 + 2) + 31) / 32 * 32;     // + 2 guard states past the end
    if (threads > 1024) threads = 1024;
    size_t smem = 2 * (size_t)(2 * T_text + 1 + 4) * sizeof(float);
    S2S_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), st));
    const int SH = ((2 * T_text + 1 + 2) + 31) / 32 * 32;
    if (dlogp && 2 * SH <= 1024) {                 // alpha and beta sweeps side by side in one CTA
        forward_sum_par_kernel<<<B, 2 * SH, 4 * (size_t)(2 * T_text + 1 + 4) * sizeof(float), st>>>(
            logp, prior, text_lens, feats_lens, B, T_feats, T_text, SH, blank_logp, alpha_ws, loss, dlogp, grad_scale);
        S2S_LAUNCH_OK();
        return S2S_OK;
    }
    forward_sum_kernel<<<B, threads, smem, st>>>(logp, prior, text_lens, feats_lens, B, T_feats, T_text, blank_logp, alpha_ws, loss,
                                                 dlogp, grad_scale);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_gauss_weights(const float* ds, const int32_t* feats_lens, const int32_t* text_lens, void* P, int B, int T_feats,
                                 int T_text, int64_t ldP, float delta, int dtype, void* stream) {
    S2S_REQUIRE(ds && feats_lens && text_lens && P && B > 0 && T_feats > 0 && T_text > 0 && ldP >= T_text && B <= 65535,
                "gauss_weights: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div_l(T_feats, 8), (unsigned)B);
    S2S_DISPATCH_DTYPE(dtype, T, (gauss_weights_kernel<T><<<grid, 256, (size_t)T_text * sizeof(float), st>>>(
        ds, feats_lens, text_lens, (T*)P, T_feats, T_text, ldP, delta)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_duration_loss(const void* pre, const float* ds, const int32_t* text_lens, int B, int T_text, float offset,
                                 float clamp_max, float grad_scale, const float* g_douts, float* d_outs, float* loss, void* d_pre,
                                 int dtype, void* stream) {
    S2S_REQUIRE(pre && ds && text_lens && B > 0 && T_text > 0, "duration_loss: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    S2S_DISPATCH_DTYPE(dtype, T, (duration_loss_kernel<T><<<1, 256, 0, st>>>((const T*)pre, ds, text_lens, B, T_text, offset, clamp_max,
                                                                            grad_scale, g_douts, d_outs, loss, (T*)d_pre)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_duration_infer(const void* pre, float* d, int n, float offset, float clamp_max, int dtype, void* stream) {
    S2S_REQUIRE(pre && d && n > 0, "duration_infer: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    S2S_DISPATCH_DTYPE(dtype, T, (duration_infer_kernel<T><<<(unsigned)ceil_div_l(n, 256), 256, 0, st>>>((const T*)pre, d, n, offset, clamp_max)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
