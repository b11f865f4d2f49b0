// LengthRegulator (modules/length_regulator.py:46-97 of the reference): expand (B, T, D) to frame level by repeating row i of
// utterance b ds[b, i] times, pad the ragged result with pad_value.  Per-utterance exclusive prefix sums of the durations turn
// the repeat into a gather (forward: one binary search per output frame) and its adjoint into a run sum per source row, so
// neither direction needs atomics and both read / write every element once with 16-byte accesses when D allows.
#include "common.cuh"

namespace s2s {

// cum[b, 0..T] = exclusive prefix sums of round(ds[b, :] * alpha) (alpha == 1: ds as is; torch.round = half to even),
// all_ones: every duration is 1 (the reference's "all predicted durations are 0" rescue, length_regulator.py:86-94)
__global__ void lr_cumsum_kernel(const long long* __restrict__ ds, int* __restrict__ cum, int T, float alpha, int all_ones) {
    const int b = blockIdx.x, lane = threadIdx.x;      // one warp per utterance, chunks of 32 durations
    const long long* d = ds + (size_t)b * T;
    int* c = cum + (size_t)b * (T + 1);
    int carry = 0;
    for (int i0 = 0; i0 < T; i0 += 32) {
        const int i = i0 + lane;
        int v = 0;
        if (i < T) {
            if (all_ones) v = 1;
            else if (alpha == 1.0f) v = (int)max(d[i], 0LL);
            else v = max((int)rintf((float)d[i] * alpha), 0);
        }
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (i < T) c[i] = carry + incl - v;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) c[T] = carry;
}

template <typename T, int VEC>
__global__ void lr_fwd_kernel(const T* __restrict__ x, const int* __restrict__ cum, T* __restrict__ y, int B, int Tin, int Lmax, int D,
                              float pad_value) {
    const int per_row = D / VEC;
    const long total = (long)B * Lmax * per_row;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int v = (int)(e % per_row);
        const long row = e / per_row;
        const int t = (int)(row % Lmax), b = (int)(row / Lmax);
        const int* c = cum + (size_t)b * (Tin + 1);
        T* dst = y + (size_t)row * D + (size_t)v * VEC;
        if (t >= c[Tin]) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) dst[q] = from_f<T>(pad_value);
            continue;
        }
        int lo = 0, hi = Tin;                           // largest i with c[i] <= t (zero-length rows are skipped by the search)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (c[mid] <= t) lo = mid; else hi = mid;
        }
        const T* src = x + ((size_t)b * Tin + lo) * D + (size_t)v * VEC;
        if (VEC == 8) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        else if (VEC == 4 && sizeof(T) == 4) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        else {
#pragma unroll
            for (int q = 0; q < VEC; ++q) dst[q] = src[q];
        }
    }
}

// dx[b, i, :] = sum of dy[b, t, :] over the run t in [cum[i], cum[i+1])
template <typename T>
__global__ void lr_bwd_kernel(const T* __restrict__ dy, const int* __restrict__ cum, T* __restrict__ dx, int B, int Tin, int Lmax, int D) {
    const long total = (long)B * Tin * D;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(e % D);
        const long row = e / D;
        const int i = (int)(row % Tin), b = (int)(row / Tin);
        const int* c = cum + (size_t)b * (Tin + 1);
        const int t0 = c[i], t1 = min(c[i + 1], Lmax);
        float acc = 0.f;
        for (int t = t0; t < t1; ++t) acc += to_f(dy[((size_t)b * Lmax + t) * D + ch]);
        dx[e] = from_f<T>(acc);
    }
}

}  // namespace s2s

using namespace s2s;

extern "C" int s2s_lr_cumsum(const int64_t* ds, int32_t* cum, int B, int T, float alpha, int all_ones, void* stream) {
    S2S_REQUIRE(ds && cum && B > 0 && T > 0 && alpha > 0.f, "lr_cumsum: bad arguments");
    lr_cumsum_kernel<<<B, 32, 0, (cudaStream_t)stream>>>((const long long*)ds, cum, T, alpha, all_ones);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_lr_fwd(const void* x, const int32_t* cum, void* y, int B, int T, int Lmax, int D, float pad_value, int dtype,
                          void* stream) {
    S2S_REQUIRE(x && cum && y && B > 0 && T > 0 && Lmax > 0 && D > 0, "lr_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int esz = dtype == S2S_BF16 ? 2 : 4, vec = 16 / esz;
    const bool wide = D % vec == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0;
    const long n = (long)B * Lmax * (wide ? D / vec : D);
    long grid = ceil_div_l(n, 256);
    if (grid > (long)num_sms() * 16) grid = (long)num_sms() * 16;
    if (dtype == S2S_F32) {
        if (wide) lr_fwd_kernel<float, 4><<<(unsigned)grid, 256, 0, st>>>((const float*)x, cum, (float*)y, B, T, Lmax, D, pad_value);
        else lr_fwd_kernel<float, 1><<<(unsigned)grid, 256, 0, st>>>((const float*)x, cum, (float*)y, B, T, Lmax, D, pad_value);
    } else if (dtype == S2S_BF16) {
        if (wide) lr_fwd_kernel<bf16, 8><<<(unsigned)grid, 256, 0, st>>>((const bf16*)x, cum, (bf16*)y, B, T, Lmax, D, pad_value);
        else lr_fwd_kernel<bf16, 1><<<(unsigned)grid, 256, 0, st>>>((const bf16*)x, cum, (bf16*)y, B, T, Lmax, D, pad_value);
    } else {
        return set_error(S2S_ERR_INVALID, "bad dtype %d", dtype);
    }
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_lr_bwd(const void* dy, const int32_t* cum, void* dx, int B, int T, int Lmax, int D, int dtype, void* stream) {
    S2S_REQUIRE(dy && cum && dx && B > 0 && T > 0 && Lmax > 0 && D > 0, "lr_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    long grid = ceil_div_l((long)B * T * D, 256);
    if (grid > (long)num_sms() * 16) grid = (long)num_sms() * 16;
    S2S_DISPATCH_DTYPE(dtype, T_, (lr_bwd_kernel<T_><<<(unsigned)grid, 256, 0, st>>>((const T_*)dy, cum, (T_*)dx, B, T, Lmax, D)));
    S2S_LAUNCH_OK();
    return S2S_OK;
}
