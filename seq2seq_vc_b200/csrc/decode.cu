// Single-position decode kernels for autoregressive inference with a KV cache (reference: VTN.inference, models/vtn.py:302-394,
// Decoder.forward_one_step, modules/transformer/decoder.py:239-273 -- whose "cache" re-projects K/V of the whole prefix every
// step; here K/V rows are projected once and kept).  Batch 1: every op is a GEMV or an attention over the cached keys, the
// position lives in a DEVICE scalar so that one captured CUDA graph serves every step.
#include "common.cuh"

namespace s2s {

// y[n] = act(sum_k W[n,k] x[k] + bias[n]) (* dropout) (+ residual[n]);  one warp per output row, x staged in shared memory
template <typename T>
__global__ void __launch_bounds__(256) gemv_kernel(const T* __restrict__ W, const float* __restrict__ bias, const T* __restrict__ x,
                                                   const T* __restrict__ residual, T* __restrict__ y, int N, int K, int relu,
                                                   Dropout drop, const int32_t* __restrict__ pos_dev) {
    extern __shared__ float xs[];
    dropout_resolve(drop);
    for (int k = threadIdx.x; k < K; k += blockDim.x) xs[k] = to_f<T>(x[k]);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    const T* w = W + (long)n * K;
    float acc = 0.f;
    if ((K & 7) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
        for (int k = lane * 8; k < K; k += 256) {
            float v[8];
            Vec8<T>::load(w + k, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = fmaf(v[i], xs[k + i], acc);
        }
    } else {
        for (int k = lane; k < K; k += 32) acc = fmaf(to_f<T>(w[k]), xs[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        float v = acc + (bias ? bias[n] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        const long pos = pos_dev ? (long)*pos_dev : 0;
        v *= dropout_factor(drop, (uint64_t)(pos * N + n));
        if (residual) v += to_f<T>(residual[n]);
        y[n] = from_f<T>(v);
    }
}

// One CTA per head.  Optionally appends this step's key / value (knew / vnew, head-major (H, dk)) to the cache at row `pos`,
// then ctx[h] = softmax_s(scale * q_h . K[s, h]) V[s, h] over s < S, S = fixed_S (source attention) or pos + 1 (self attention).
// Cache element (s, h, j) at cache + s * row_stride + h * dk + j.  probs (H, ldp) float32 optionally receives the weights.
template <typename T>
__global__ void __launch_bounds__(256) decode_attn_kernel(const T* __restrict__ q, const T* __restrict__ knew, const T* __restrict__ vnew,
                                                          T* __restrict__ kcache, T* __restrict__ vcache, long row_stride, int dk, int fixed_S,
                                                          int S_cap, const int32_t* __restrict__ pos_dev, float scale, T* __restrict__ ctx,
                                                          float* __restrict__ probs, int ldp, long probs_step_stride) {
    extern __shared__ float sm[];           // [S_cap] scores, then [dk] q, then [256] reduction scratch
    float* sc = sm;
    float* qs = sm + S_cap;
    float* red = qs + dk;
    const int h = blockIdx.x, tid = threadIdx.x;
    const int pos = pos_dev ? *pos_dev : 0;
    int S = fixed_S >= 0 ? fixed_S : pos + 1;
    if (S > S_cap) S = S_cap;
    if (knew && tid < dk) {
        kcache[(long)pos * row_stride + h * dk + tid] = knew[h * dk + tid];
        vcache[(long)pos * row_stride + h * dk + tid] = vnew[h * dk + tid];
    }
    if (tid < dk) qs[tid] = to_f<T>(q[h * dk + tid]);
    __syncthreads();
    // scores (thread per key)
    float mx = -INFINITY;
    for (int s = tid; s < S; s += blockDim.x) {
        const T* kr = kcache + (long)s * row_stride + h * dk;
        float d = 0.f;
        for (int j = 0; j < dk; ++j) d = fmaf(qs[j], to_f<T>(kr[j]), d);
        d *= scale;
        sc[s] = d;
        mx = fmaxf(mx, d);
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = -INFINITY;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int s = tid; s < S; s += blockDim.x) {
        const float e = __expf(sc[s] - mx);
        sc[s] = e;
        sum += e;
    }
    sum = block_sum(sum, red);
    const float inv = S > 0 ? 1.f / sum : 0.f;
    __syncthreads();
    if (probs)
        for (int s = tid; s < ldp; s += blockDim.x) probs[(long)pos * probs_step_stride + (long)h * ldp + s] = s < S ? sc[s] * inv : 0.f;
    // ctx[j] = sum_s p[s] V[s][j]: thread (grp, j) sums keys s = grp (mod ngrp)
    const int ngrp = blockDim.x / dk;
    const int j = tid % dk, grp = tid / dk;
    float acc = 0.f;
    if (grp < ngrp)
        for (int s = grp; s < S; s += ngrp) acc = fmaf(sc[s], to_f<T>(vcache[(long)s * row_stride + h * dk + j]), acc);
    __syncthreads();
    red[tid] = (grp < ngrp) ? acc : 0.f;
    __syncthreads();
    if (tid < dk) {
        float v = 0.f;
        for (int g2 = 0; g2 < ngrp; ++g2) v += red[g2 * dk + tid];
        ctx[h * dk + tid] = from_f<T>(v * inv);
    }
}

// y = x + alpha * pe[pos]  (ScaledPositionalEncoding for the single new position; eval mode: no dropout)
template <typename T>
__global__ void decode_pe_kernel(const T* __restrict__ x, const float* __restrict__ pe, const float* __restrict__ alpha,
                                 const int32_t* __restrict__ pos_dev, T* __restrict__ y, int d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d) y[i] = from_f<T>(to_f<T>(x[i]) + alpha[0] * pe[(long)(*pos_dev) * d + i]);
}

// end of a decode step: the next decoder input is the last of the r generated frames; also records frames / logits of the step
template <typename T>
__global__ void decode_advance_kernel(const T* __restrict__ feat, const T* __restrict__ logit, T* __restrict__ next_in,
                                      float* __restrict__ frames, float* __restrict__ logits, int32_t* __restrict__ pos_dev, int odim, int r) {
    const int pos = *pos_dev;
    for (int i = threadIdx.x; i < odim * r; i += blockDim.x) frames[(long)pos * odim * r + i] = to_f<T>(feat[i]);
    for (int i = threadIdx.x; i < r; i += blockDim.x) logits[(long)pos * r + i] = to_f<T>(logit[i]);
    for (int i = threadIdx.x; i < odim; i += blockDim.x) next_in[i] = feat[(r - 1) * odim + i];
    __syncthreads();
    if (threadIdx.x == 0) *pos_dev = pos + 1;
}

}  // namespace s2s

using namespace s2s;

extern "C" int s2s_gemv(const void* W, const float* bias, const void* x, const void* residual, void* y, int N, int K, int relu,
                        const s2s_dropout_t* drop, const int32_t* pos_dev, int dtype, void* stream) {
    S2S_REQUIRE(W && x && y && N > 0 && K > 0 && K <= 12288, "gemv: bad arguments (K <= 12288)");
    cudaStream_t st = (cudaStream_t)stream;
    Dropout d = make_dropout(drop);
    S2S_DISPATCH_DTYPE(dtype, T, (gemv_kernel<T><<<(unsigned)ceil_div_l(N, 8), 256, (size_t)K * sizeof(float), st>>>(
        (const T*)W, bias, (const T*)x, (const T*)residual, (T*)y, N, K, relu, d, pos_dev)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_decode_attn(const void* q, const void* knew, const void* vnew, void* kcache, void* vcache, int64_t row_stride, int H,
                               int dk, int fixed_S, int S_cap, const int32_t* pos_dev, float scale, void* ctx, float* probs, int ldp,
                               int64_t probs_step_stride, int dtype, void* stream) {
    S2S_REQUIRE(q && kcache && vcache && ctx && H > 0 && dk > 0 && dk <= 256 && S_cap > 0, "decode_attn: bad arguments (d_k <= 256)");
    S2S_REQUIRE((knew == nullptr) == (vnew == nullptr), "decode_attn: knew and vnew come together");
    S2S_REQUIRE(fixed_S >= 0 || pos_dev != nullptr, "decode_attn: self-attention needs the device position");
    S2S_REQUIRE(probs == nullptr || ldp > 0, "decode_attn: probs needs its row length");
    const size_t smem = ((size_t)S_cap + dk + 256) * sizeof(float);
    S2S_REQUIRE(smem <= 48 * 1024, "decode_attn: S_cap %d too large for one CTA (<= ~11k keys)", S_cap);
    cudaStream_t st = (cudaStream_t)stream;
    S2S_DISPATCH_DTYPE(dtype, T, (decode_attn_kernel<T><<<H, 256, smem, st>>>((const T*)q, (const T*)knew, (const T*)vnew, (T*)kcache,
                                                                               (T*)vcache, row_stride, dk, fixed_S, S_cap, pos_dev, scale,
                                                                               (T*)ctx, probs, ldp, (long)probs_step_stride)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_decode_pe(const void* x, const float* pe, const float* alpha, const int32_t* pos_dev, void* y, int d, int dtype,
                             void* stream) {
    S2S_REQUIRE(x && pe && alpha && pos_dev && y && d > 0, "decode_pe: bad arguments");
    S2S_DISPATCH_DTYPE(dtype, T, (decode_pe_kernel<T><<<(unsigned)ceil_div_l(d, 128), 128, 0, (cudaStream_t)stream>>>(
        (const T*)x, pe, alpha, pos_dev, (T*)y, d)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_decode_advance(const void* feat, const void* logit, void* next_in, float* frames, float* logits, int32_t* pos_dev,
                                  int odim, int r, int dtype, void* stream) {
    S2S_REQUIRE(feat && logit && next_in && frames && logits && pos_dev && odim > 0 && r > 0, "decode_advance: bad arguments");
    S2S_DISPATCH_DTYPE(dtype, T, (decode_advance_kernel<T><<<1, 128, 0, (cudaStream_t)stream>>>((const T*)feat, (const T*)logit, (T*)next_in,
                                                                                                 frames, logits, pos_dev, odim, r)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
