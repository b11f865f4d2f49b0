// Masked softmax over attention scores (forward and backward), one warp per (b, h, query) row.
// Mirrors modules/transformer/attention.py:76-85 of the reference: masked keys are excluded
// from the normalisation and come out as exact zeros; rows with no visible key are all zeros.
#include "common.cuh"

namespace s2s {

template <typename T> struct RowIO {  // 4-wide when possible
    static __device__ __forceinline__ void ld4(const T* p, float (&v)[4]) { Vec4<T>::load(p, v); }
    static __device__ __forceinline__ void st4(T* p, const float (&v)[4]) { Vec4<T>::store(p, v); }
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(128) softmax_fwd_kernel(const T* __restrict__ S, T* __restrict__ P,
                                                          T* __restrict__ Pd, const int32_t* __restrict__ klens,
                                                          int B, int H, int T1, int T2, long ld, int causal,
                                                          Dropout drop) {
    dropout_resolve(drop);
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const long rows = (long)B * H * T1;
    if (row >= rows) return;
    const int i = (int)(row % T1);
    const int b = (int)(row / ((long)H * T1));
    int limit = klens ? klens[b] : T2;
    if (limit > T2) limit = T2;
    if (causal && limit > i + 1) limit = i + 1;
    if (limit < 0) limit = 0;
    const T* s = S + row * ld;
    T* p = P + row * ld;
    T* pd = Pd ? Pd + row * ld : nullptr;
    const int ldi = (int)ld;

    float mx = -INFINITY;
    if (VEC) {
        for (int c = lane * 4; c < limit; c += 128) {
            float v[4];
            RowIO<T>::ld4(s + c, v);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c + k < limit) mx = fmaxf(mx, v[k]);
        }
    } else {
        for (int c = lane; c < limit; c += 32) mx = fmaxf(mx, to_f<T>(s[c]));
    }
    mx = warp_max(mx);
    float sum = 0.f;
    if (VEC) {
        for (int c = lane * 4; c < limit; c += 128) {
            float v[4];
            RowIO<T>::ld4(s + c, v);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c + k < limit) sum += __expf(v[k] - mx);
        }
    } else {
        for (int c = lane; c < limit; c += 32) sum += __expf(to_f<T>(s[c]) - mx);
    }
    sum = warp_sum(sum);
    const float inv = (limit > 0) ? 1.f / sum : 0.f;
    const uint64_t base = (uint64_t)row * (uint64_t)T2;
    if (VEC) {
        for (int c = lane * 4; c < ldi; c += 128) {
            float v[4], o[4], od[4];
            RowIO<T>::ld4(s + c, v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[k] = (c + k < limit) ? __expf(v[k] - mx) * inv : 0.f;
                od[k] = pd ? o[k] * dropout_factor(drop, base + c + k) : 0.f;
            }
            RowIO<T>::st4(p + c, o);
            if (pd) RowIO<T>::st4(pd + c, od);
        }
    } else {
        for (int c = lane; c < ldi; c += 32) {
            float o = (c < limit) ? __expf(to_f<T>(s[c]) - mx) * inv : 0.f;
            p[c] = from_f<T>(o);
            if (pd) pd[c] = from_f<T>(o * dropout_factor(drop, base + c));
        }
    }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(128) softmax_bwd_kernel(const T* __restrict__ P, T* __restrict__ dP, long rows,
                                                          int T2, long ld, float scale, Dropout drop) {
    dropout_resolve(drop);
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* p = P + row * ld;
    T* g = dP + row * ld;
    const uint64_t base = (uint64_t)row * (uint64_t)T2;
    const int ldi = (int)ld;
    float dot = 0.f;
    if (VEC) {
        for (int c = lane * 4; c < T2; c += 128) {
            float a[4], d[4];
            RowIO<T>::ld4(p + c, a);
            RowIO<T>::ld4(g + c, d);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c + k < T2) dot += a[k] * d[k] * dropout_factor(drop, base + c + k);
        }
    } else {
        for (int c = lane; c < T2; c += 32) dot += to_f<T>(p[c]) * to_f<T>(g[c]) * dropout_factor(drop, base + c);
    }
    dot = warp_sum(dot);
    if (VEC) {
        for (int c = lane * 4; c < ldi; c += 128) {
            float a[4], d[4], o[4];
            RowIO<T>::ld4(p + c, a);
            RowIO<T>::ld4(g + c, d);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = (c + k < T2) ? scale * a[k] * (d[k] * dropout_factor(drop, base + c + k) - dot) : 0.f;
            RowIO<T>::st4(g + c, o);
        }
    } else {
        for (int c = lane; c < ldi; c += 32) {
            float o = (c < T2) ? scale * to_f<T>(p[c]) * (to_f<T>(g[c]) * dropout_factor(drop, base + c) - dot) : 0.f;
            g[c] = from_f<T>(o);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Row-in-registers variants (ld % 8 == 0, ld <= 1024): each lane keeps NV 16-byte chunks of its row, so S / P / dP are
// read from HBM exactly once and written once (the generic kernels above re-read the row per pass).
// ---------------------------------------------------------------------------------------------
template <typename T, int NV>
__global__ void __launch_bounds__(256) softmax_fwd_reg_kernel(const T* __restrict__ S, T* __restrict__ P, T* __restrict__ Pd,
                                                              const int32_t* __restrict__ klens, int B, int H, int T1, int T2, int ld,
                                                              int causal, Dropout drop) {
    dropout_resolve(drop);
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long rows = (long)B * H * T1;
    if (row >= rows) return;
    const int i = (int)(row % T1);
    const int b = (int)(row / ((long)H * T1));
    int limit = klens ? klens[b] : T2;
    if (limit > T2) limit = T2;
    if (causal && limit > i + 1) limit = i + 1;
    if (limit < 0) limit = 0;
    const T* s = S + row * ld;
    float v[NV][8];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < ld) {
            Vec8<T>::load(s + c, v[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c + k < limit) mx = fmaxf(mx, v[j][k]);
        }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float e = (c + k < limit) ? __expf(v[j][k] - mx) : 0.f;
            v[j][k] = e;
            sum += e;
        }
    }
    sum = warp_sum(sum);
    const float inv = (limit > 0) ? 1.f / sum : 0.f;
    const uint64_t base = (uint64_t)row * (uint64_t)T2;
    const bool even = ((T2 & 7) == 0);      // dropout_factors<8> needs an 8-aligned element index
    T* p = P + row * ld;
    T* pd = Pd ? Pd + row * ld : nullptr;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < ld) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = v[j][k] * inv;
            Vec8<T>::store(p + c, o);
            if (pd) {
                float m[8];
                if (even) dropout_factors<8>(drop, base + c, m);
                else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) m[k] = dropout_factor(drop, base + c + k);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] *= m[k];
                Vec8<T>::store(pd + c, o);
            }
        }
    }
}

template <typename T, int NV>
__global__ void __launch_bounds__(256) softmax_bwd_reg_kernel(const T* __restrict__ P, T* __restrict__ dP, long rows, int T2, int ld,
                                                              float scale, Dropout drop) {
    dropout_resolve(drop);
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const T* p = P + row * ld;
    T* g = dP + row * ld;
    const uint64_t base = (uint64_t)row * (uint64_t)T2;
    const bool even = ((T2 & 7) == 0);      // dropout_factors<8> needs an 8-aligned element index
    float a[NV][8], d[NV][8];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < ld) {
            Vec8<T>::load(p + c, a[j]);
            Vec8<T>::load(g + c, d[j]);
            float m[8];
            if (even) dropout_factors<8>(drop, base + c, m);
            else {
#pragma unroll
                for (int k = 0; k < 8; ++k) m[k] = dropout_factor(drop, base + c + k);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                d[j][k] = (c + k < T2) ? d[j][k] * m[k] : 0.f;
                if (c + k >= T2) a[j][k] = 0.f;
                dot += a[j][k] * d[j][k];
            }
        }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int c = (j * 32 + lane) * 8;
        if (c < ld) {
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = scale * a[j][k] * (d[j][k] - dot);
            Vec8<T>::store(g + c, o);
        }
    }
}

}  // namespace s2s

using namespace s2s;

static bool rows_vec_ok(int64_t ld, const void* a, const void* b, const void* c, int dtype) {
    size_t es = dtype == S2S_F32 ? 4 : 2;
    auto al = [&](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % (4 * es)) == 0; };
    return (ld % 4 == 0) && al(a) && al(b) && al(c);
}

extern "C" int s2s_softmax_fwd(const void* S, void* P, void* Pd, const int32_t* klens, int B, int H, int T1, int T2,
                               int64_t ld, int causal, const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(S && P && B > 0 && H > 0 && T1 > 0 && T2 > 0 && ld >= T2, "softmax_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    Dropout d = make_dropout(drop);
    long rows = (long)B * H * T1;
    if (ld % 8 == 0 && ld <= 1024 && aligned16(S, P, Pd)) {
        const int nv = (int)((ld + 255) / 256);
        unsigned g8 = (unsigned)ceil_div_l(rows, 8);
        S2S_DISPATCH_DTYPE(dtype, T, {
            if (nv == 1) softmax_fwd_reg_kernel<T, 1><<<g8, 256, 0, st>>>((const T*)S, (T*)P, (T*)Pd, klens, B, H, T1, T2, (int)ld, causal, d);
            else if (nv == 2) softmax_fwd_reg_kernel<T, 2><<<g8, 256, 0, st>>>((const T*)S, (T*)P, (T*)Pd, klens, B, H, T1, T2, (int)ld, causal, d);
            else if (nv == 3) softmax_fwd_reg_kernel<T, 3><<<g8, 256, 0, st>>>((const T*)S, (T*)P, (T*)Pd, klens, B, H, T1, T2, (int)ld, causal, d);
            else softmax_fwd_reg_kernel<T, 4><<<g8, 256, 0, st>>>((const T*)S, (T*)P, (T*)Pd, klens, B, H, T1, T2, (int)ld, causal, d);
        });
        S2S_LAUNCH_OK();
        return S2S_OK;
    }
    bool ok = rows_vec_ok(ld, S, P, Pd, dtype);
    unsigned grid = (unsigned)ceil_div_l(rows, 4);
    S2S_DISPATCH_DTYPE(dtype, T, {
        if (ok) softmax_fwd_kernel<T, true><<<grid, 128, 0, st>>>((const T*)S, (T*)P, (T*)Pd, klens, B, H, T1, T2, ld, causal, d);
        else softmax_fwd_kernel<T, false><<<grid, 128, 0, st>>>((const T*)S, (T*)P, (T*)Pd, klens, B, H, T1, T2, ld, causal, d);
    });
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_softmax_bwd(const void* P, void* dP, int B, int H, int T1, int T2, int64_t ld, float scale,
                               const s2s_dropout_t* drop, int dtype, void* stream) {
    S2S_REQUIRE(P && dP && B > 0 && H > 0 && T1 > 0 && T2 > 0 && ld >= T2, "softmax_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    Dropout d = make_dropout(drop);
    long rows = (long)B * H * T1;
    if (ld % 8 == 0 && ld <= 1024 && aligned16(P, dP)) {
        const int nv = (int)((ld + 255) / 256);
        unsigned g8 = (unsigned)ceil_div_l(rows, 8);
        S2S_DISPATCH_DTYPE(dtype, T, {
            if (nv == 1) softmax_bwd_reg_kernel<T, 1><<<g8, 256, 0, st>>>((const T*)P, (T*)dP, rows, T2, (int)ld, scale, d);
            else if (nv == 2) softmax_bwd_reg_kernel<T, 2><<<g8, 256, 0, st>>>((const T*)P, (T*)dP, rows, T2, (int)ld, scale, d);
            else if (nv == 3) softmax_bwd_reg_kernel<T, 3><<<g8, 256, 0, st>>>((const T*)P, (T*)dP, rows, T2, (int)ld, scale, d);
            else softmax_bwd_reg_kernel<T, 4><<<g8, 256, 0, st>>>((const T*)P, (T*)dP, rows, T2, (int)ld, scale, d);
        });
        S2S_LAUNCH_OK();
        return S2S_OK;
    }
    bool ok = rows_vec_ok(ld, P, dP, nullptr, dtype);
    unsigned grid = (unsigned)ceil_div_l(rows, 4);
    S2S_DISPATCH_DTYPE(dtype, T, {
        if (ok) softmax_bwd_kernel<T, true><<<grid, 128, 0, st>>>((const T*)P, (T*)dP, rows, T2, ld, scale, d);
        else softmax_bwd_kernel<T, false><<<grid, 128, 0, st>>>((const T*)P, (T*)dP, rows, T2, ld, scale, d);
    });
    S2S_LAUNCH_OK();
    return S2S_OK;
}
