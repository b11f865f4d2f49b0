// Stochastic duration predictor of AAS-VC (the shipped recipe's default, egs/arctic/vc2/conf/aas_vc.melmelmel.v1.yaml:57):
// the element-wise / per-position pieces of
//   StochasticDurationPredictor.forward          (modules/duration_predictor.py:131-304)
//   DilatedDepthSeparableConv, ConvFlow, ElementwiseAffineFlow, LogFlow   (modules/vits/flow.py:19-310)
//   piecewise rational-quadratic spline with linear tails                (modules/vits/transform.py:12-216)
// The 1x1 convolutions are s2s_gemm calls and the channel LayerNorms s2s_layernorm_* over the channels-last (B, T, C) layout;
// what is here: exact GELU, the dilated depthwise convolution, the spline (forward, inverse, and its gradient by forward-mode
// dual numbers: 30 inputs per element, evaluated direction by direction -- the tensors are (B, T_text) small, arithmetic is
// free, and the softmax -> cumsum -> bin search -> rational-quadratic chain stays one readable function for both passes),
// the flow heads (affine, sigmoid / log change of variables, Gaussian log-likelihoods), a counter-based normal generator,
// and duration read-out.  Everything is float32: logs, square roots and 1e-3 floors do not survive bf16.
// Noise is an explicit input of the flows (the reference draws it with torch.randn inside forward).
#include "common.cuh"

namespace s2s {
namespace sdp {

constexpr int BINS = 10, NPAR = 3 * BINS - 1;
constexpr float TAIL = 5.0f, MIN_W = 1e-3f, MIN_H = 1e-3f, MIN_D = 1e-3f;

// ---- forward-mode dual number -----------------------------------------------------------------
struct Dual {
    float v, d;
};
__device__ __forceinline__ Dual mk(float v, float d = 0.f) { return Dual{v, d}; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
    const float q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
__device__ __forceinline__ Dual operator*(float s, Dual a) { return {s * a.v, s * a.d}; }
__device__ __forceinline__ Dual operator+(Dual a, float s) { return {a.v + s, a.d}; }
__device__ __forceinline__ Dual operator-(Dual a, float s) { return {a.v - s, a.d}; }
__device__ __forceinline__ Dual dexp(Dual a) { const float e = expf(a.v); return {e, e * a.d}; }
__device__ __forceinline__ Dual dlog(Dual a) { return {logf(a.v), a.d / a.v}; }
__device__ __forceinline__ Dual dsqrt(Dual a) { const float s = sqrtf(a.v); return {s, 0.5f * a.d / s}; }
__device__ __forceinline__ Dual dsoftplus(Dual a) {          // F.softplus: threshold 20
    if (a.v > 20.f) return a;
    const float e = expf(a.v);
    return {log1pf(e), a.d * e / (1.f + e)};
}
// plain floats through the same template
__device__ __forceinline__ float mk_f(float v) { return v; }
__device__ __forceinline__ float dexp(float a) { return expf(a); }
__device__ __forceinline__ float dlog(float a) { return logf(a); }
__device__ __forceinline__ float dsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float dsoftplus(float a) { return a > 20.f ? a : log1pf(expf(a)); }
__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(Dual a) { return a.v; }
template <typename S> __device__ __forceinline__ S lift(float v);
template <> __device__ __forceinline__ float lift<float>(float v) { return v; }
template <> __device__ __forceinline__ Dual lift<Dual>(float v) { return Dual{v, 0.f}; }

// softmax -> floor -> cumulative knots on [-TAIL, TAIL] with pinned ends (transform.py:117-135): cum[0..BINS]
template <typename S>
__device__ __forceinline__ void knots(const S (&u)[BINS], float min_size, S (&cum)[BINS + 1]) {
    float mx = val(u[0]);
#pragma unroll
    for (int i = 1; i < BINS; ++i) mx = fmaxf(mx, val(u[i]));
    S e[BINS];
    S sum = lift<S>(0.f);
#pragma unroll
    for (int i = 0; i < BINS; ++i) { e[i] = dexp(u[i] - mx); sum = sum + e[i]; }
    S run = lift<S>(0.f);
    cum[0] = lift<S>(-TAIL);
#pragma unroll
    for (int i = 0; i < BINS; ++i) {
        const S w = (1.f - min_size * BINS) * (e[i] / sum) + min_size;
        run = run + w;
        cum[i + 1] = (2.f * TAIL) * run - TAIL;
    }
    cum[BINS] = lift<S>(TAIL);
}

// rq spline with linear tails (transform.py:44-209).  h: [BINS widths | BINS heights | BINS - 1 derivatives] (widths / heights
// already divided by sqrt(hidden)).  Returns y and log|det| (negated for the inverse, as the reference does).
template <typename S>
__device__ void rq_spline(S x, const S (&uw)[BINS], const S (&uh)[BINS], const S (&ud)[BINS - 1], bool inverse, S& y, S& lad) {
    const float xv = val(x);
    if (!(xv >= -TAIL && xv <= TAIL)) { y = x; lad = lift<S>(0.f); return; }
    S cumw[BINS + 1], cumh[BINS + 1];
    knots(uw, MIN_W, cumw);
    knots(uh, MIN_H, cumh);
    // bin search on the value parts (transform.py:212-216: the last knot is nudged by 1e-6)
    int idx = -1;
#pragma unroll
    for (int i = 0; i <= BINS; ++i) {
        const float loc = val(inverse ? cumh[i] : cumw[i]) + (i == BINS ? 1e-6f : 0.f);
        idx += (xv >= loc) ? 1 : 0;
    }
    idx = idx < 0 ? 0 : (idx > BINS - 1 ? BINS - 1 : idx);
    const float cboundary = logf(expf(1.f - MIN_D) - 1.f);           // boundary derivative = 1 (transform.py:66-68)
    S in_cumw = cumw[0], in_w = cumw[1] - cumw[0], in_cumh = cumh[0], in_h = cumh[1] - cumh[0];
    S d0 = lift<S>(0.f), d1 = lift<S>(0.f);
#pragma unroll
    for (int i = 0; i < BINS; ++i) {
        if (i == idx) {
            in_cumw = cumw[i]; in_w = cumw[i + 1] - cumw[i];
            in_cumh = cumh[i]; in_h = cumh[i + 1] - cumh[i];
            const S u0 = (i == 0) ? lift<S>(cboundary) : ud[i == 0 ? 0 : i - 1];
            const S u1 = (i == BINS - 1) ? lift<S>(cboundary) : ud[i == BINS - 1 ? BINS - 2 : i];
            d0 = dsoftplus(u0) + MIN_D;
            d1 = dsoftplus(u1) + MIN_D;
        }
    }
    const S delta = in_h / in_w;
    S th;
    if (inverse) {
        const S t = (x - in_cumh) * (d0 + d1 - 2.f * delta);
        const S a = t + in_h * (delta - d0);
        const S b = in_h * d0 - t;
        const S c = -(delta * (x - in_cumh));
        th = (2.f * c) / (-b - dsqrt(b * b - 4.f * (a * c)));
        y = th * in_w + in_cumw;
    } else {
        th = (x - in_cumw) / in_w;
    }
    const S one_m = lift<S>(1.f) - th;
    const S tt = th * one_m;
    const S denom = delta + (d0 + d1 - 2.f * delta) * tt;
    if (!inverse) y = in_cumh + in_h * (delta * th * th + d0 * tt) / denom;
    const S num = delta * delta * (d1 * th * th + 2.f * (delta * tt) + d0 * one_m * one_m);
    lad = dlog(num) - 2.f * dlog(denom);
    if (inverse) lad = -lad;
}

// y / lad of element n (masked positions: y = 0, lad = 0 -- conv_flow multiplies both by the mask)
__global__ void __launch_bounds__(128) spline_fwd_kernel(const float* __restrict__ x, long x_bs, const float* __restrict__ h,
                                                         const int32_t* __restrict__ tlens, float* __restrict__ y, long y_bs,
                                                         float* __restrict__ lad, int B, int T, float inv_den, int inverse) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B * T) return;
    const int b = n / T, t = n - b * T;
    if (t >= tlens[b]) { y[b * y_bs + t] = 0.f; if (lad) lad[n] = 0.f; return; }
    const float* hp = h + (long)n * NPAR;
    float uw[BINS], uh[BINS], ud[BINS - 1];
#pragma unroll
    for (int i = 0; i < BINS; ++i) { uw[i] = hp[i] * inv_den; uh[i] = hp[BINS + i] * inv_den; }
#pragma unroll
    for (int i = 0; i < BINS - 1; ++i) ud[i] = hp[2 * BINS + i];
    float yy, ll;
    rq_spline<float>(x[b * x_bs + t], uw, uh, ud, inverse != 0, yy, ll);
    y[b * y_bs + t] = yy;
    if (lad) lad[n] = ll;
}

// gradient of sum(gy * y + glad * lad) w.r.t. x and the 29 spline parameters: one dual evaluation per input direction --
// one WARP per element, lane j evaluates direction j (lanes 0..28: the parameters, lane 29: the abscissa)
__global__ void __launch_bounds__(128) spline_bwd_kernel(const float* __restrict__ x, long x_bs, const float* __restrict__ h,
                                                         const int32_t* __restrict__ tlens, const float* __restrict__ gy, long gy_bs,
                                                         const float* __restrict__ glad, float* __restrict__ dx, long dx_bs,
                                                         float* __restrict__ dh, int B, int T, float inv_den) {
    const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int j = threadIdx.x & 31;
    if (n >= B * T) return;
    const int b = n / T, t = n - b * T;
    float* dhp = dh + (long)n * NPAR;
    if (t >= tlens[b]) {
        if (j < NPAR) dhp[j] = 0.f;
        if (j == NPAR) dx[b * dx_bs + t] = 0.f;
        return;
    }
    if (j > NPAR) return;
    const float* hp = h + (long)n * NPAR;
    const float g_y = gy ? gy[b * gy_bs + t] : 0.f, g_l = glad ? glad[n] : 0.f;
    const float xv = x[b * x_bs + t];
    Dual uw[BINS], uh[BINS], ud[BINS - 1];
#pragma unroll
    for (int i = 0; i < BINS; ++i) {
        uw[i] = Dual{hp[i] * inv_den, j == i ? inv_den : 0.f};
        uh[i] = Dual{hp[BINS + i] * inv_den, j == BINS + i ? inv_den : 0.f};
    }
#pragma unroll
    for (int i = 0; i < BINS - 1; ++i) ud[i] = Dual{hp[2 * BINS + i], j == 2 * BINS + i ? 1.f : 0.f};
    Dual yy, ll;
    rq_spline<Dual>(Dual{xv, j == NPAR ? 1.f : 0.f}, uw, uh, ud, false, yy, ll);
    const float g = g_y * yy.d + g_l * ll.d;
    if (j == NPAR) dx[b * dx_bs + t] = g; else dhp[j] = g;
}

// ---- exact GELU ----------------------------------------------------------------------------------
__device__ __forceinline__ float gelu1(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu1(float v) {
    return 0.5f * (1.f + erff(v * 0.70710678118654752f)) + v * 0.39894228040143268f * expf(-0.5f * v * v);
}
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long n) {
    const long n4 = n >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        reinterpret_cast<float4*>(y)[i] = make_float4(gelu1(v.x), gelu1(v.y), gelu1(v.z), gelu1(v.w));
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) y[i] = gelu1(x[i]);
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, long n) {
    const long n4 = n >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(x)[i], g = reinterpret_cast<const float4*>(dy)[i];
        reinterpret_cast<float4*>(dx)[i] = make_float4(g.x * dgelu1(v.x), g.y * dgelu1(v.y), g.z * dgelu1(v.z), g.w * dgelu1(v.w));
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) dx[i] = dy[i] * dgelu1(x[i]);
}

// ---- Conv1d(1 -> C, k = 1): y[n, c] = x[n] w[c] + b[c]  (ConvFlow.input_conv, post_pre) -------------------------------------
__global__ void __launch_bounds__(128) outer_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                        float* __restrict__ y, long N, int C) {
    for (long n = blockIdx.x; n < N; n += gridDim.x) {
        const float xv = x[n];
        for (int c = threadIdx.x; c < C; c += blockDim.x) y[n * C + c] = fmaf(xv, w[c], b[c]);
    }
}
// dx[n] = sum_c dy[n, c] w[c]
__global__ void __launch_bounds__(128) outer_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, long N, int C) {
    __shared__ float red[32];
    for (long n = blockIdx.x; n < N; n += gridDim.x) {
        float s = 0.f;
        for (int c = threadIdx.x; c < C; c += blockDim.x) s = fmaf(dy[n * C + c], w[c], s);
        s = block_sum(s, red);
        if (threadIdx.x == 0) dx[n] = s;
        __syncthreads();
    }
}
// dw[c] += sum_n dy[n, c] x[n] ; db[c] += sum_n dy[n, c].  blockDim = 128 channels, gridDim.y row chunks
__global__ void __launch_bounds__(128) outer_bwd_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                                                           float* __restrict__ db, long N, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const long per = (N + gridDim.y - 1) / gridDim.y;
    const long r0 = blockIdx.y * per, r1 = (r0 + per < N) ? r0 + per : N;
    float aw = 0.f, ab = 0.f;
    for (long r = r0; r < r1; ++r) {
        const float g = dy[r * C + c];
        aw = fmaf(g, x[r], aw);
        ab += g;
    }
    atomicAdd(dw + c, aw);
    atomicAdd(db + c, ab);
}

// ---- dilated depthwise Conv1d over time on the masked input (flow.py:192-204), channels-last ---------
// one block per (b, t) row, each thread owns VEC consecutive channels (16-byte loads when C % 4 == 0): no per-element division
template <int VEC>
__global__ void __launch_bounds__(128) dwdil_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ tlens, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ y, int B, int T, int C, int K, int dil) {
    const int half = (K - 1) / 2;
    for (long r = blockIdx.x; r < (long)B * T; r += gridDim.x) {
        const int b = (int)(r / T), t = (int)(r - (long)b * T);
        const int tl = min(tlens[b], T);
        for (int c = threadIdx.x * VEC; c < C; c += blockDim.x * VEC) {
            float acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = bias[c + e];
            for (int j = 0; j < K; ++j) {
                const int s = t + (j - half) * dil;
                if (s >= 0 && s < tl) {
                    float v[VEC];
                    if (VEC == 4) Vec4<float>::load(x + ((long)b * T + s) * C + c, reinterpret_cast<float(&)[4]>(v)); else v[0] = x[((long)b * T + s) * C + c];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[e] = fmaf(w[(c + e) * K + j], v[e], acc[e]);
                }
            }
            if (VEC == 4) Vec4<float>::store(y + r * C + c, reinterpret_cast<float(&)[4]>(acc)); else y[r * C + c] = acc[0];
        }
    }
}
// dx[b,s,c] = mask(s) * sum_j w[c,j] dy[b, s - (j - half) dil, c]
template <int VEC>
__global__ void __launch_bounds__(128) dwdil_bwd_dx_kernel(const float* __restrict__ dy, const int32_t* __restrict__ tlens, const float* __restrict__ w,
                                                           float* __restrict__ dx, int B, int T, int C, int K, int dil) {
    const int half = (K - 1) / 2;
    for (long r = blockIdx.x; r < (long)B * T; r += gridDim.x) {
        const int b = (int)(r / T), s = (int)(r - (long)b * T);
        const bool on = s < tlens[b];
        for (int c = threadIdx.x * VEC; c < C; c += blockDim.x * VEC) {
            float acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
            if (on) {
                for (int j = 0; j < K; ++j) {
                    const int t = s - (j - half) * dil;
                    if (t >= 0 && t < T) {
                        float v[VEC];
                        if (VEC == 4) Vec4<float>::load(dy + ((long)b * T + t) * C + c, reinterpret_cast<float(&)[4]>(v)); else v[0] = dy[((long)b * T + t) * C + c];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[e] = fmaf(w[(c + e) * K + j], v[e], acc[e]);
                    }
                }
            }
            if (VEC == 4) Vec4<float>::store(dx + r * C + c, reinterpret_cast<float(&)[4]>(acc)); else dx[r * C + c] = acc[0];
        }
    }
}
// dw[c,j] += sum_{b,t} dy[b,t,c] xm[b, t + (j - half) dil, c] ; db[c] += sum dy.  blockDim = 128 channels, gridDim.y row chunks
template <int K>
__global__ void dwdil_bwd_dw_kernel(const float* __restrict__ dy, const float* __restrict__ x, const int32_t* __restrict__ tlens,
                                    float* __restrict__ dw, float* __restrict__ db, int B, int T, int C, int dil) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    constexpr int half = (K - 1) / 2;
    float aw[K], ab = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) aw[j] = 0.f;
    const long rows = (long)B * T;
    const long per = (rows + gridDim.y - 1) / gridDim.y;
    const long r0 = blockIdx.y * per, r1 = (r0 + per < rows) ? r0 + per : rows;
    for (long r = r0; r < r1; ++r) {
        const int b = (int)(r / T), t = (int)(r % T);
        const int tl = tlens[b];
        const float g = dy[r * C + c];
        ab += g;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int s = t + (j - half) * dil;
            if (s >= 0 && s < T && s < tl) aw[j] = fmaf(g, x[((long)b * T + s) * C + c], aw[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < K; ++j) atomicAdd(dw + c * K + j, aw[j]);
    atomicAdd(db + c, ab);
}

// ---- flow heads on z (B, 2, T) -------------------------------------------------------------------
__device__ __forceinline__ float logsigmoidf(float v) { return fminf(v, 0.f) - log1pf(expf(-fabsf(v))); }

// ElementwiseAffineFlow (flow.py:96-112): y = (m + exp(logs) x) mask; logdet_b = sum_t mask (logs_0 + logs_1).
// inverse: y = (x - m) exp(-logs) mask.  nll[b] (optional) += sign * logdet_b.
__global__ void affine_fwd_kernel(const float* __restrict__ z, const float* __restrict__ m, const float* __restrict__ logs,
                                  const int32_t* __restrict__ tlens, float* __restrict__ y, float* __restrict__ nll, float sign, int B, int T,
                                  int inverse) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B * T) return;
    const int b = n / T, t = n - b * T;
    const bool on = t < tlens[b];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const float x = z[((long)b * 2 + ch) * T + t];
        float o = 0.f;
        if (on) o = inverse ? (x - m[ch]) * expf(-logs[ch]) : m[ch] + expf(logs[ch]) * x;
        y[((long)b * 2 + ch) * T + t] = o;
    }
    if (nll && t == 0) atomicAdd(nll + b, sign * (logs[0] + logs[1]) * (float)min(tlens[b], T));
}
// dz, dm, dlogs of the forward direction; g_nll[b] = d loss / d nll[b]
__global__ void affine_bwd_kernel(const float* __restrict__ z, const float* __restrict__ logs, const int32_t* __restrict__ tlens,
                                  const float* __restrict__ gy, const float* __restrict__ g_nll, float sign, float* __restrict__ dz,
                                  float* __restrict__ dm, float* __restrict__ dlogs, int B, int T) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    float am[2] = {0.f, 0.f}, al[2] = {0.f, 0.f};
    if (n < B * T) {
        const int b = n / T, t = n - b * T;
        const bool on = t < tlens[b];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const long i = ((long)b * 2 + ch) * T + t;
            const float g = on ? gy[i] : 0.f, e = expf(logs[ch]);
            dz[i] = g * e;
            am[ch] = g;
            al[ch] = g * e * z[i] + (on && g_nll ? sign * g_nll[b] : 0.f);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const float sm = warp_sum(am[ch]), sl = warp_sum(al[ch]);
        if ((threadIdx.x & 31) == 0) { atomicAdd(dm + ch, sm); atomicAdd(dlogs + ch, sl); }
    }
}

// Posterior head (duration_predictor.py:262-279): from z_q = (z_u, z1) and the durations w
//   u = sigmoid(z_u) mask; z0 = (w - u) mask; y0 = log(max(z0, 1e-5)) mask          (LogFlow, flow.py:62-65)
//   out = (y0, z1);  nll[b] += sum_t mask [ (logsigmoid(z_u) + logsigmoid(-z_u))  (that is -logdet_q ... see below)  + y0 ]
// sign bookkeeping: nll + logq = ... - logdet_tot + (-0.5 e^2 terms) - logdet_tot_q, where logdet_tot gets sum(-y0) and
// logdet_tot_q gets sum(logsigmoid(z_u) + logsigmoid(-z_u)): the head contributes  + sum(y0) - sum(lsig)  to nll.
__global__ void head_fwd_kernel(const float* __restrict__ zq, const float* __restrict__ w, const int32_t* __restrict__ tlens,
                                float* __restrict__ out, float* __restrict__ nll, int B, int T) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    float contrib = 0.f;
    int b = 0;
    if (n < B * T) {
        b = n / T;
        const int t = n - b * T;
        const bool on = t < tlens[b];
        const float zu = zq[((long)b * 2) * T + t], z1 = zq[((long)b * 2 + 1) * T + t];
        float y0 = 0.f;
        if (on) {
            const float u = 1.f / (1.f + expf(-zu));
            const float z0 = w[n] - u;
            y0 = logf(fmaxf(z0, 1e-5f));
            contrib = y0 - (logsigmoidf(zu) + logsigmoidf(-zu));
        }
        out[((long)b * 2) * T + t] = y0;
        out[((long)b * 2 + 1) * T + t] = on ? z1 : 0.f;
        atomicAdd(nll + b, contrib);
    }
}
__global__ void head_bwd_kernel(const float* __restrict__ zq, const float* __restrict__ w, const int32_t* __restrict__ tlens,
                                const float* __restrict__ gout, const float* __restrict__ g_nll, float* __restrict__ dzq, int B, int T) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B * T) return;
    const int b = n / T, t = n - b * T;
    const bool on = t < tlens[b];
    const long i0 = ((long)b * 2) * T + t, i1 = i0 + T;
    float d0 = 0.f, d1 = 0.f;
    if (on) {
        const float zu = zq[i0];
        const float u = 1.f / (1.f + expf(-zu));
        const float z0 = w[n] - u;
        const float gn = g_nll[b];
        // y0 = log(max(z0, eps)): d y0 / d zu = -(u (1 - u)) / z0 when z0 > eps, else 0
        const float dy0 = (z0 > 1e-5f) ? -(u * (1.f - u)) / z0 : 0.f;
        // d (logsigmoid(zu) + logsigmoid(-zu)) / d zu = (1 - u) - u
        d0 = (gout[i0] + gn) * dy0 - gn * (1.f - 2.f * u);
        d1 = gout[i1];
    }
    dzq[i0] = d0;
    dzq[i1] = d1;
}

// Gaussian terms: nll[b] += sign * sum_{ch,t} mask * 0.5 * (log 2 pi + z^2); gradient: dz = sign * g_nll[b] * z * mask
__global__ void gauss_fwd_kernel(const float* __restrict__ z, const int32_t* __restrict__ tlens, float* __restrict__ nll, float sign, int B, int T) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B * T) return;
    const int b = n / T, t = n - b * T;
    if (t >= tlens[b]) return;
    const float a = z[((long)b * 2) * T + t], c = z[((long)b * 2 + 1) * T + t];
    atomicAdd(nll + b, sign * (1.8378770664093453f + 0.5f * (a * a + c * c)));
}
__global__ void gauss_bwd_kernel(const float* __restrict__ z, const int32_t* __restrict__ tlens, const float* __restrict__ g_nll, float sign,
                                 float* __restrict__ dz, int accumulate, int B, int T) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B * T) return;
    const int b = n / T, t = n - b * T;
    const bool on = t < tlens[b];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const long i = ((long)b * 2 + ch) * T + t;
        const float g = on ? sign * g_nll[b] * z[i] : 0.f;
        dz[i] = accumulate ? dz[i] + g : g;
    }
}
// nll[b] += sign * sum_t x[b, t]  (already masked per-position terms: the log-determinants of the splines)
__global__ void rowsum_kernel(const float* __restrict__ x, float* __restrict__ nll, float sign, int B, int T) {
    const int b = blockIdx.x;
    float s = 0.f;
    for (int t = threadIdx.x; t < T; t += blockDim.x) s += x[(long)b * T + t];
    __shared__ float red[32];
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(nll + b, sign * s);
}
// out[b, t] = sign * g[b]  (gradient of the row sum)
__global__ void rowbcast_kernel(const float* __restrict__ g, float* __restrict__ out, float sign, int B, int T) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < B * T) out[n] = sign * g[n / T];
}

// standard normal draws: counter-based (murmur-mixed index) Box-Muller, one pair per thread; seed_dev makes CUDA-graph replays differ
__global__ void randn_kernel(float* __restrict__ out, long n, uint64_t seed, const uint64_t* __restrict__ seed_dev, uint64_t stream_id) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i >= n) return;
    uint64_t key = seed * 0x9e3779b97f4a7c15ull + stream_id * 0xd1b54a32d192ed03ull + 0x2545f4914f6cdd1dull;
    if (seed_dev) key += (*seed_dev) * 0x9e3779b97f4a7c15ull;
    uint32_t a = mix32((uint32_t)i * 0x9e3779b1u + (uint32_t)key);
    a = mix32(a ^ (uint32_t)(key >> 32));
    uint32_t b = mix32(a + 0x85ebca6bu + (uint32_t)(i >> 32));
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777217.0f);      // (0, 1)
    const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);               // [0, 1)
    const float r = sqrtf(-2.f * logf(u1));
    float s, c;
    sincospif(2.f * u2, &s, &c);
    out[2 * i] = r * c;
    if (2 * i + 1 < n) out[2 * i + 1] = r * s;
}

// dur[b, t] = min(ceil(exp(z[b, 0, t]) * mask), clamp_max)     (duration_predictor.py:298-304, aas_vc.py:393)
__global__ void dur_kernel(const float* __restrict__ z, const int32_t* __restrict__ tlens, float* __restrict__ dur, float clamp_max, int B, int T) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= B * T) return;
    const int b = n / T, t = n - b * T;
    const float w = (t < tlens[b]) ? expf(z[((long)b * 2) * T + t]) : 0.f;
    dur[n] = fminf(ceilf(w), clamp_max);
}

}  // namespace sdp
}  // namespace s2s

using namespace s2s;
using namespace s2s::sdp;

static inline unsigned nblk(long n, int per) { return (unsigned)ceil_div_l(n, per); }

extern "C" int s2s_gelu_fwd(const float* x, float* y, int64_t n, void* stream) {
    S2S_REQUIRE(x && y, "gelu_fwd: null pointer");
    if (n <= 0) return S2S_OK;
    gelu_fwd_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_gelu_bwd(const float* dy, const float* x, float* dx, int64_t n, void* stream) {
    S2S_REQUIRE(dy && x && dx, "gelu_bwd: null pointer");
    if (n <= 0) return S2S_OK;
    gelu_bwd_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, x, dx, n);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_outer_fwd(const float* x, const float* w, const float* b, float* y, int64_t N, int C, void* stream) {
    S2S_REQUIRE(x && w && b && y && N > 0 && C > 0, "outer_fwd: bad arguments");
    outer_fwd_kernel<<<ew_grid(N, 1), 128, 0, (cudaStream_t)stream>>>(x, w, b, y, N, C);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_outer_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int64_t N, int C, void* stream) {
    S2S_REQUIRE(dy && x && w && N > 0 && C > 0, "outer_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dx) {
        outer_bwd_dx_kernel<<<ew_grid(N, 1), 128, 0, st>>>(dy, w, dx, N, C);
        S2S_LAUNCH_OK();
    }
    if (dw && db) {
        long chunks = ceil_div_l(N, 64);
        const long cap = (long)num_sms() * 4 / ceil_div_l(C, 128);
        if (chunks > cap) chunks = cap < 1 ? 1 : cap;
        outer_bwd_dw_kernel<<<dim3((unsigned)ceil_div_l(C, 128), (unsigned)chunks), 128, 0, st>>>(dy, x, dw, db, N, C);
        S2S_LAUNCH_OK();
    }
    return S2S_OK;
}
extern "C" int s2s_dwconv_dilated_fwd(const float* x, const int32_t* tlens, const float* w, const float* bias, float* y, int B, int T, int C,
                                      int K, int dil, void* stream) {
    S2S_REQUIRE(x && tlens && w && bias && y && B > 0 && T > 0 && C > 0 && K >= 1 && (K & 1) && dil >= 1, "dwconv_dilated_fwd: bad arguments");
    if (C % 4 == 0 && aligned16(x, y)) dwdil_fwd_kernel<4><<<ew_grid((long)B * T, 1), 128, 0, (cudaStream_t)stream>>>(x, tlens, w, bias, y, B, T, C, K, dil);
    else dwdil_fwd_kernel<1><<<ew_grid((long)B * T, 1), 128, 0, (cudaStream_t)stream>>>(x, tlens, w, bias, y, B, T, C, K, dil);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_dwconv_dilated_bwd(const float* dy, const float* x, const int32_t* tlens, const float* w, float* dx, float* dw, float* db,
                                      int B, int T, int C, int K, int dil, void* stream) {
    S2S_REQUIRE(dy && x && tlens && w && B > 0 && T > 0 && C > 0 && (K == 3 || K == 5 || K == 7) && dil >= 1, "dwconv_dilated_bwd: bad arguments (K in 3, 5, 7)");
    cudaStream_t st = (cudaStream_t)stream;
    if (dx) {
        if (C % 4 == 0 && aligned16(dy, dx)) dwdil_bwd_dx_kernel<4><<<ew_grid((long)B * T, 1), 128, 0, st>>>(dy, tlens, w, dx, B, T, C, K, dil);
        else dwdil_bwd_dx_kernel<1><<<ew_grid((long)B * T, 1), 128, 0, st>>>(dy, tlens, w, dx, B, T, C, K, dil);
        S2S_LAUNCH_OK();
    }
    if (dw && db) {
        long chunks = ceil_div_l((long)B * T, 64);
        const long cap = (long)num_sms() * 4 / ceil_div_l(C, 128);
        if (chunks > cap) chunks = cap < 1 ? 1 : cap;
        dim3 grid((unsigned)ceil_div_l(C, 128), (unsigned)chunks);
        if (K == 3) dwdil_bwd_dw_kernel<3><<<grid, 128, 0, st>>>(dy, x, tlens, dw, db, B, T, C, dil);
        else if (K == 5) dwdil_bwd_dw_kernel<5><<<grid, 128, 0, st>>>(dy, x, tlens, dw, db, B, T, C, dil);
        else dwdil_bwd_dw_kernel<7><<<grid, 128, 0, st>>>(dy, x, tlens, dw, db, B, T, C, dil);
        S2S_LAUNCH_OK();
    }
    return S2S_OK;
}
extern "C" int s2s_rq_spline_fwd(const float* x, int64_t x_bs, const float* h, const int32_t* tlens, float* y, int64_t y_bs, float* lad, int B,
                                 int T, float hidden, int inverse, void* stream) {
    S2S_REQUIRE(x && h && tlens && y && B > 0 && T > 0 && hidden > 0.f, "rq_spline_fwd: bad arguments");
    spline_fwd_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(x, x_bs, h, tlens, y, y_bs, lad, B, T, 1.f / sqrtf(hidden), inverse);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_rq_spline_bwd(const float* x, int64_t x_bs, const float* h, const int32_t* tlens, const float* gy, int64_t gy_bs,
                                 const float* glad, float* dx, int64_t dx_bs, float* dh, int B, int T, float hidden, void* stream) {
    S2S_REQUIRE(x && h && tlens && dx && dh && B > 0 && T > 0 && hidden > 0.f, "rq_spline_bwd: bad arguments");
    spline_bwd_kernel<<<nblk((long)B * T * 32, 128), 128, 0, (cudaStream_t)stream>>>(x, x_bs, h, tlens, gy, gy_bs, glad, dx, dx_bs, dh, B, T,
                                                                                      1.f / sqrtf(hidden));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_sdp_affine_fwd(const float* z, const float* m, const float* logs, const int32_t* tlens, float* y, float* nll, float sign,
                                  int B, int T, int inverse, void* stream) {
    S2S_REQUIRE(z && m && logs && tlens && y && B > 0 && T > 0, "sdp_affine_fwd: bad arguments");
    affine_fwd_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(z, m, logs, tlens, y, nll, sign, B, T, inverse);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_sdp_affine_bwd(const float* z, const float* logs, const int32_t* tlens, const float* gy, const float* g_nll, float sign,
                                  float* dz, float* dm, float* dlogs, int B, int T, void* stream) {
    S2S_REQUIRE(z && logs && tlens && gy && dz && dm && dlogs && B > 0 && T > 0, "sdp_affine_bwd: bad arguments");
    affine_bwd_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(z, logs, tlens, gy, g_nll, sign, dz, dm, dlogs, B, T);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_sdp_head_fwd(const float* zq, const float* w, const int32_t* tlens, float* out, float* nll, int B, int T, void* stream) {
    S2S_REQUIRE(zq && w && tlens && out && nll && B > 0 && T > 0, "sdp_head_fwd: bad arguments");
    head_fwd_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(zq, w, tlens, out, nll, B, T);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_sdp_head_bwd(const float* zq, const float* w, const int32_t* tlens, const float* gout, const float* g_nll, float* dzq, int B,
                                int T, void* stream) {
    S2S_REQUIRE(zq && w && tlens && gout && g_nll && dzq && B > 0 && T > 0, "sdp_head_bwd: bad arguments");
    head_bwd_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(zq, w, tlens, gout, g_nll, dzq, B, T);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_sdp_gauss_fwd(const float* z, const int32_t* tlens, float* nll, float sign, int B, int T, void* stream) {
    S2S_REQUIRE(z && tlens && nll && B > 0 && T > 0, "sdp_gauss_fwd: bad arguments");
    gauss_fwd_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(z, tlens, nll, sign, B, T);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_sdp_gauss_bwd(const float* z, const int32_t* tlens, const float* g_nll, float sign, float* dz, int accumulate, int B, int T,
                                 void* stream) {
    S2S_REQUIRE(z && tlens && g_nll && dz && B > 0 && T > 0, "sdp_gauss_bwd: bad arguments");
    gauss_bwd_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(z, tlens, g_nll, sign, dz, accumulate, B, T);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_rowsum_acc(const float* x, float* acc, float sign, int B, int T, void* stream) {
    S2S_REQUIRE(x && acc && B > 0 && T > 0, "rowsum_acc: bad arguments");
    rowsum_kernel<<<(unsigned)B, 128, 0, (cudaStream_t)stream>>>(x, acc, sign, B, T);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_rowbcast(const float* g, float* out, float sign, int B, int T, void* stream) {
    S2S_REQUIRE(g && out && B > 0 && T > 0, "rowbcast: bad arguments");
    rowbcast_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(g, out, sign, B, T);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_randn(float* out, int64_t n, uint64_t seed, const uint64_t* seed_dev, uint64_t stream_id, void* stream) {
    S2S_REQUIRE(out && n > 0, "randn: bad arguments");
    randn_kernel<<<nblk((n + 1) / 2, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, seed_dev, stream_id);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
extern "C" int s2s_sdp_durations(const float* z, const int32_t* tlens, float* dur, float clamp_max, int B, int T, void* stream) {
    S2S_REQUIRE(z && tlens && dur && B > 0 && T > 0, "sdp_durations: bad arguments");
    dur_kernel<<<nblk((long)B * T, 128), 128, 0, (cudaStream_t)stream>>>(z, tlens, dur, clamp_max, B, T);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
