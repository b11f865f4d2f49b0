// STFT -> log-mel front end (bin/preprocess.py:30-92 of the reference, librosa semantics).
//
// Fast path (n_fft = 2048): ONE WARP PER FRAME, no block-level synchronisation.
//   The real 2048-point FFT is a 1024-point complex FFT of z[m] = x[2m] + i x[2m+1], computed as
//   32 x 32 Cooley-Tukey: every lane runs a 32-point FFT in registers over the stride-32 samples it
//   loaded straight from HBM (reflect padding and the window folded into the load), applies the
//   inter-stage twiddle by recurrence, the warp transposes through its private 8 KB of shared
//   memory (__syncwarp only), every lane runs a second 32-point FFT, and the spectrum is unpacked
//   to |X[k]|, k = 0..1024.  The mel projection walks each band's non-zero bins (triangular
//   filters are sparse: ~2 x 1025 non-zeros for 80 bands) and the 80 log values are written coalesced.
//   Arithmetic: ~64 kFLOP per frame in fp32; each sample is re-used by n_fft / hop ~ 6.8 frames, so
//   the kernel is bound by fp32 FFT arithmetic, not by its 1.52 kB / frame of HBM traffic.
// Generic path (other power-of-two n_fft): one CTA per frame, radix-2 in shared memory.
#include "common.cuh"

namespace s2s {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// e^{-2 pi i j / 32}, j = 0..15
__device__ constexpr float kCos32[16] = {1.0f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f, 0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f};
__device__ constexpr float kSin32[16] = {0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f, -1.0f, -0.98078528f, -0.923879533f, -0.831469612f, -0.707106781f, -0.555570233f, -0.382683432f, -0.195090322f};

__host__ __device__ constexpr int brev5(int x) {
    return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// in-place 32-point forward DFT, decimation in frequency, fully unrolled; output v[brev5(k)] = X[k]
__device__ __forceinline__ void fft32_dif(float2 (&v)[32]) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = 16 >> s;              // butterfly span
#pragma unroll
        for (int g = 0; g < 32; g += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const float2 a = v[g + j], b = v[g + j + half];
                v[g + j] = make_float2(a.x + b.x, a.y + b.y);
                const float2 d = make_float2(a.x - b.x, a.y - b.y);
                const int tw = j << s;         // W_32^(j * 2^s)
                if (tw == 0) v[g + j + half] = d;
                else if (tw == 8) v[g + j + half] = make_float2(d.y, -d.x);   // * (-i)
                else v[g + j + half] = make_float2(d.x * kCos32[tw] - d.y * kSin32[tw], d.x * kSin32[tw] + d.y * kCos32[tw]);
            }
        }
    }
}

constexpr int LM_WARPS = 12;
constexpr int LM_ZLD = 33;                        // padded row length of the per-warp transpose / spectrum buffer

struct LogmelShared {
    // offsets (in floats) inside dynamic shared memory for the fast path
    int tw1024, tw2048, win, wpk, rng, per_warp, warp_stride;
};

__global__ void __launch_bounds__(LM_WARPS * 32) logmel2048_kernel(const float* __restrict__ wav, const float* __restrict__ window,
                                                                  const float* __restrict__ basis, float* __restrict__ mel, int B,
                                                                  int ns, int hop, int n_frames, int n_mels, float eps,
                                                                  float log_scale, int wpk_cap, const float* __restrict__ nmean,
                                                                  const float* __restrict__ nscale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NFFT = 2048, H = 1024, NB = 1025;
    float2* tw1024 = reinterpret_cast<float2*>(smem_raw);           // e^{-2 pi i j / 1024}, j < 32 only needed as base; keep 32
    float2* tw2048 = tw1024 + 32;                                   // e^{-2 pi i k / 2048}, k <= 1024
    float* swin = reinterpret_cast<float*>(tw2048 + NB + 1);        // [2048]
    float* wpk = swin + NFFT;                                       // packed non-zero filter weights [wpk_cap]
    int* rng = reinterpret_cast<int*>(wpk + wpk_cap);               // [n_mels][3] = start, end, packed offset
    float* warp_base = reinterpret_cast<float*>(rng + ((3 * n_mels + 4) & ~3));   // keep 16-byte alignment
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // per-warp scratch: zbuf float2[32 * 33] (transpose, then spectrum), mag float[1025 + pad]
    float2* zbuf = reinterpret_cast<float2*>(warp_base + (size_t)warp * (2 * 32 * LM_ZLD + NB + 7));
    float* mag = reinterpret_cast<float*>(zbuf + 32 * LM_ZLD);

    // ---- CTA-wide tables (built once per persistent CTA)
    if (tid < 32) {
        float s, c;
        sincospif(-2.0f * (float)tid / 1024.f, &s, &c);
        tw1024[tid] = make_float2(c, s);
    }
    for (int k = tid; k <= H; k += blockDim.x) {
        float s, c;
        sincospif(-2.0f * (float)k / 2048.f, &s, &c);
        tw2048[k] = make_float2(c, s);
    }
    for (int k = tid; k < NFFT; k += blockDim.x) swin[k] = window[k];
    for (int m = warp; m < n_mels; m += LM_WARPS) {
        int lo = NB, hi = -1;
        for (int k = lane; k < NB; k += 32)
            if (basis[(size_t)m * NB + k] != 0.f) { lo = min(lo, k); hi = max(hi, k); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { rng[3 * m] = lo; rng[3 * m + 1] = hi + 1; }
    }
    __syncthreads();
    if (tid == 0) {
        int off = 0;
        for (int m = 0; m < n_mels; ++m) {
            int len = max(rng[3 * m + 1] - rng[3 * m], 0);
            if (off + len > wpk_cap) { len = 0; rng[3 * m + 1] = rng[3 * m]; }
            rng[3 * m + 2] = off;
            off += len;
        }
    }
    __syncthreads();
    for (int m = warp; m < n_mels; m += LM_WARPS) {
        const int lo = rng[3 * m], hi = rng[3 * m + 1], off = rng[3 * m + 2];
        for (int k = lo + lane; k < hi; k += 32) wpk[off + k - lo] = basis[(size_t)m * NB + k];
    }
    __syncthreads();

    const long total = (long)B * n_frames;
    const long wstride = (long)gridDim.x * LM_WARPS;
    for (long fr = (long)blockIdx.x * LM_WARPS + warp; fr < total; fr += wstride) {
        const int b = (int)(fr / n_frames), f = (int)(fr % n_frames);
        const float* x = wav + (size_t)b * ns;
        const long start = (long)f * hop - H;                       // center = True: frame f covers [f*hop - n_fft/2, ...)
        // ---- step 1: lane n2 loads z[32 n1 + n2] = (x[2m], x[2m+1]) * window, n1 = 0..31 (coalesced 256 B per n1)
        float2 v[32];
        const bool interior = (start >= 0) && (start + NFFT <= ns);
        const bool even = (((size_t)b * ns + start) & 1) == 0 && ((reinterpret_cast<uintptr_t>(wav) & 7) == 0);
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int mi = 32 * n1 + lane;
            long s0 = start + 2 * mi, s1 = s0 + 1;
            if (!interior) {
                if (s0 < 0) s0 = -s0; else if (s0 >= ns) s0 = 2L * (ns - 1) - s0;
                if (s1 < 0) s1 = -s1; else if (s1 >= ns) s1 = 2L * (ns - 1) - s1;
            }
            const float2 w = *reinterpret_cast<const float2*>(swin + 2 * mi);
            if (interior && even) {
                const float2 xv = *reinterpret_cast<const float2*>(x + s0);
                v[n1] = make_float2(xv.x * w.x, xv.y * w.y);
            } else {
                v[n1] = make_float2(x[s0] * w.x, x[s1] * w.y);
            }
        }
        fft32_dif(v);                                               // v[brev5(k1)] = sum_n1 z[32 n1 + lane] W_32^(n1 k1)
        // ---- step 2: twiddle W_1024^(lane * k1) by recurrence, step 3: transpose through shared memory
        {
            const float2 wl = tw1024[lane];
            float2 w = make_float2(1.f, 0.f);
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) {
                const float2 y = cmul(v[brev5(k1)], w);
                zbuf[k1 * LM_ZLD + lane] = y;                       // row k1, column n2 = lane
                w = cmul(w, wl);
            }
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) v[n2] = zbuf[lane * LM_ZLD + n2];    // lane = k1 now
        __syncwarp();
        fft32_dif(v);                                               // v[brev5(k2)] = Z[k1 + 32 k2]
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) zbuf[k2 * LM_ZLD + lane] = v[brev5(k2)];   // Z[k] at [(k >> 5) * 33 + (k & 31)]
        __syncwarp();
        // ---- real-FFT unpack: X[k] = (Z[k] + conj Z[H-k]) / 2 - i/2 * W_2048^k (Z[k] - conj Z[H-k]), k = 0..1024
        for (int k = lane; k <= H; k += 32) {
            const int ka = k & (H - 1), kb = (H - k) & (H - 1);
            const float2 zk = zbuf[(ka >> 5) * LM_ZLD + (ka & 31)];
            float2 zc = zbuf[(kb >> 5) * LM_ZLD + (kb & 31)];
            zc.y = -zc.y;
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            const float2 o = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
            const float2 ow = cmul(o, tw2048[k]);                   // then multiply by -i: (a + ib)(-i) = b - ia
            const float re = e.x + ow.y, im = e.y - ow.x;
            mag[k] = sqrtf(re * re + im * im);
        }
        __syncwarp();
        // ---- mel bands: 8-lane groups, 4 bands per pass (triangular filters are narrow at low frequencies), + log
        float* out = mel + (size_t)fr * n_mels;
        const int grp = lane >> 3, gl = lane & 7;
        for (int m0 = 0; m0 < n_mels; m0 += 4) {
            const int m = m0 + grp;
            float acc = 0.f;
            if (m < n_mels) {
                const int lo = rng[3 * m], hi = rng[3 * m + 1], off = rng[3 * m + 2];
                for (int k = lo + gl; k < hi; k += 8) acc = fmaf(mag[k], wpk[off + k - lo], acc);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (gl == 0 && m < n_mels) {
                float v = log2f(fmaxf(eps, acc)) * log_scale;
                if (nmean) v = (v - nmean[m]) / nscale[m];          // StandardScaler.transform (bin/normalize.py:193) fused
                out[m] = v;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// generic power-of-two path: one CTA per frame, radix-2 in shared memory
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) logmel_kernel(const float* __restrict__ wav, const float* __restrict__ window,
                                                     const float* __restrict__ basis, float* __restrict__ mel, int B,
                                                     int ns, int n_fft, int log2h, int hop, int n_frames, int n_mels,
                                                     float eps, float log_scale, const float* __restrict__ nmean,
                                                     const float* __restrict__ nscale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int half = n_fft >> 1, nbins = half + 1;
    float2* data = reinterpret_cast<float2*>(smem_raw);       // [half]
    float2* tw = data + half;                                 // [half]  e^{-2 pi i k / n_fft}
    float* mag = reinterpret_cast<float*>(tw + half);         // [nbins (+pad)]
    float* wpk = mag + nbins + 3;                             // packed non-zero filter weights [<= 4*nbins]
    int* rng = reinterpret_cast<int*>(wpk + 4 * nbins);       // [n_mels][3] = start, end, packed offset
    float* swin = reinterpret_cast<float*>(rng + 3 * n_mels); // [n_fft]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int k = tid; k < half; k += 256) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)n_fft, &s, &c);
        tw[k] = make_float2(c, s);
    }
    for (int k = tid; k < n_fft; k += 256) swin[k] = window[k];
    // non-zero range of every mel band (bands are contiguous triangles); one warp per band
    for (int m = warp; m < n_mels; m += 8) {
        int lo = nbins, hi = -1;
        for (int k = lane; k < nbins; k += 32)
            if (basis[(size_t)m * nbins + k] != 0.f) { lo = min(lo, k); hi = max(hi, k); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { rng[3 * m] = lo; rng[3 * m + 1] = hi + 1; }
    }
    __syncthreads();
    if (tid == 0) {
        int off = 0;
        for (int m = 0; m < n_mels; ++m) {
            int len = max(rng[3 * m + 1] - rng[3 * m], 0);
            if (off + len > 4 * nbins) { len = 0; rng[3 * m + 1] = rng[3 * m]; }  // cannot happen for triangular banks
            rng[3 * m + 2] = off;
            off += len;
        }
    }
    __syncthreads();
    for (int m = warp; m < n_mels; m += 8) {
        int lo = rng[3 * m], hi = rng[3 * m + 1], off = rng[3 * m + 2];
        for (int k = lo + lane; k < hi; k += 32) wpk[off + k - lo] = basis[(size_t)m * nbins + k];
    }
    __syncthreads();

    const long total = (long)B * n_frames;
    const int pad = n_fft >> 1;
    for (long fr = blockIdx.x; fr < total; fr += gridDim.x) {
        const int b = (int)(fr / n_frames), f = (int)(fr % n_frames);
        const float* x = wav + (size_t)b * ns;
        const long start = (long)f * hop - pad;
        // ---- load, window, pack even/odd samples into complex points, bit-reversed
        for (int n = tid; n < half; n += 256) {
            long s0 = start + 2 * n, s1 = s0 + 1;
            if (s0 < 0) s0 = -s0; else if (s0 >= ns) s0 = 2L * (ns - 1) - s0;
            if (s1 < 0) s1 = -s1; else if (s1 >= ns) s1 = 2L * (ns - 1) - s1;
            float a = x[s0] * swin[2 * n], c = x[s1] * swin[2 * n + 1];
            unsigned r = __brev((unsigned)n) >> (32 - log2h);
            data[r] = make_float2(a, c);
        }
        __syncthreads();
        // ---- radix-2 DIT over `half` complex points
        for (int s = 0; s < log2h; ++s) {
            const int hs = 1 << s;
            for (int t = tid; t < (half >> 1); t += 256) {
                int j = t & (hs - 1);
                int i0 = ((t >> s) << (s + 1)) + j;
                int i1 = i0 + hs;
                // twiddle e^{-2 pi i j / (2 hs)} = tw[j * n_fft / (2 hs)] = tw[j << (log2h - s)]
                float2 w = tw[j << (log2h - s)];
                float2 u = data[i0], v = cmul(data[i1], w);
                data[i0] = make_float2(u.x + v.x, u.y + v.y);
                data[i1] = make_float2(u.x - v.x, u.y - v.y);
            }
            __syncthreads();
        }
        // ---- real-FFT unpack: X[k] = (Z[k] + conj Z[h-k])/2 - i/2 * W^k (Z[k] - conj Z[h-k])
        for (int k = tid; k <= half; k += 256) {
            float2 zk = data[k & (half - 1)];
            float2 zc = data[(half - k) & (half - 1)];
            zc.y = -zc.y;
            float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            float2 o = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
            float2 w = (k < half) ? tw[k] : make_float2(-1.f, 0.f);
            float2 ow = cmul(o, w);  // then multiply by -i: (a + ib)(-i) = b - ia
            float re = e.x + ow.y, im = e.y - ow.x;
            mag[k] = sqrtf(re * re + im * im);
        }
        __syncthreads();
        // ---- mel band sums + log
        float* out = mel + (size_t)fr * n_mels;
        for (int m = warp; m < n_mels; m += 8) {
            int lo = rng[3 * m], hi = rng[3 * m + 1], off = rng[3 * m + 2];
            float acc = 0.f;
            for (int k = lo + lane; k < hi; k += 32) acc = fmaf(mag[k], wpk[off + k - lo], acc);
            acc = warp_sum(acc);
            if (lane == 0) {
                float v = log2f(fmaxf(eps, acc)) * log_scale;
                if (nmean) v = (v - nmean[m]) / nscale[m];
                out[m] = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace s2s

using namespace s2s;

static int logmel_impl(const float* wav, const float* window, const float* mel_basis, float* mel, int B, int n_samples, int n_fft,
                       int hop, int n_mels, float eps, float log_base, const float* nmean, const float* nscale, void* stream);

extern "C" int s2s_logmel(const float* wav, const float* window, const float* mel_basis, float* mel, int B,
                          int n_samples, int n_fft, int hop, int n_mels, float eps, float log_base, void* stream) {
    return logmel_impl(wav, window, mel_basis, mel, B, n_samples, n_fft, hop, n_mels, eps, log_base, nullptr, nullptr, stream);
}

extern "C" int s2s_logmel_norm(const float* wav, const float* window, const float* mel_basis, const float* mean, const float* scale,
                               float* mel, int B, int n_samples, int n_fft, int hop, int n_mels, float eps, float log_base,
                               void* stream) {
    S2S_REQUIRE(mean && scale, "logmel_norm: mean / scale are required");
    return logmel_impl(wav, window, mel_basis, mel, B, n_samples, n_fft, hop, n_mels, eps, log_base, mean, scale, stream);
}

static int logmel_impl(const float* wav, const float* window, const float* mel_basis, float* mel, int B, int n_samples, int n_fft,
                       int hop, int n_mels, float eps, float log_base, const float* nmean, const float* nscale, void* stream) {
    S2S_REQUIRE(wav && window && mel_basis && mel, "logmel: null pointer");
    S2S_REQUIRE(B > 0 && n_samples > 0 && hop > 0 && n_mels > 0, "logmel: bad shape");
    S2S_REQUIRE(n_fft >= 64 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "logmel: n_fft %d must be a power of two in [64, 4096]", n_fft);
    S2S_REQUIRE(n_samples > n_fft / 2, "logmel: reflect padding needs n_samples > n_fft/2");
    int log2n = 0;
    while ((1 << log2n) < n_fft) ++log2n;
    const int half = n_fft / 2, nbins = half + 1;
    const int n_frames = 1 + n_samples / hop;
    // log_b(x) = log2(x) / log2(b)
    double lb = (log_base == 0.f) ? 2.718281828459045 : (double)log_base;
    S2S_REQUIRE(lb > 1.0, "logmel: bad log base");
    float log_scale = (float)(1.0 / log2(lb));
    long total = (long)B * n_frames;
    if (n_fft == 2048 && n_mels <= 256) {
        // warp-per-frame fast path
        const int wpk_cap = (4 * nbins + 3) & ~3;
        size_t smem = (size_t)(32 + nbins + 1) * 8 + (size_t)n_fft * 4 + (size_t)wpk_cap * 4 + (size_t)((3 * n_mels + 4) & ~3) * 4 +
                      (size_t)LM_WARPS * (2 * 32 * LM_ZLD + nbins + 7) * 4;
        static bool attr2 = false;
        if (!attr2) {
            S2S_CUDA_OK(cudaFuncSetAttribute(logmel2048_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr2 = true;
        }
        S2S_REQUIRE(smem <= 200 * 1024, "logmel: shared memory request too large");
        long grid = (long)num_sms();
        long need = ceil_div_l(total, LM_WARPS);
        if (grid > need) grid = need;
        logmel2048_kernel<<<(unsigned)grid, LM_WARPS * 32, smem, (cudaStream_t)stream>>>(wav, window, mel_basis, mel, B, n_samples, hop,
                                                                                         n_frames, n_mels, eps, log_scale, wpk_cap, nmean, nscale);
        S2S_LAUNCH_OK();
        return S2S_OK;
    }
    size_t smem = (size_t)half * 8 * 2 + (size_t)(nbins + 3) * 4 + (size_t)4 * nbins * 4 + (size_t)3 * n_mels * 4 + (size_t)n_fft * 4;
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        S2S_CUDA_OK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    S2S_REQUIRE(smem <= 160 * 1024, "logmel: shared memory request too large");
    long grid = (long)num_sms() * 4;
    if (grid > total) grid = total;
    logmel_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(wav, window, mel_basis, mel, B, n_samples, n_fft, log2n - 1,
                                                                       hop, n_frames, n_mels, eps, log_scale, nmean, nscale);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
