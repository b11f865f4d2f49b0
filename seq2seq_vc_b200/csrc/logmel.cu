// STFT -> log-mel front end (bin/preprocess.py:30-92 of the reference, librosa semantics).
//
// Persistent CTAs; each CTA builds its twiddle table and a compressed copy of the (sparse,
// triangular) mel filterbank in shared memory once, then loops over frames:
//   reflect-padded, windowed frame -> packed as N/2 complex points (bit-reversed on store)
//   -> in-place radix-2 FFT in shared memory -> real-FFT unpack -> |X[k]| for k = 0..N/2
//   -> mel band sums over each band's non-zero bin range (warp per band, shuffle reduce)
//   -> log10(max(eps, .)).
// HBM traffic is the algorithmic minimum (every sample is read ~n_fft/hop times but the re-reads
// hit L1/L2; every output is written once); the kernel is bound by the fp32 FFT arithmetic.
#include "common.cuh"

namespace s2s {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void __launch_bounds__(256) logmel_kernel(const float* __restrict__ wav, const float* __restrict__ window,
                                                     const float* __restrict__ basis, float* __restrict__ mel, int B,
                                                     int ns, int n_fft, int log2h, int hop, int n_frames, int n_mels,
                                                     float eps, float log_scale) {
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
            if (lane == 0) out[m] = log2f(fmaxf(eps, acc)) * log_scale;
        }
        __syncthreads();
    }
}

}  // namespace s2s

using namespace s2s;

extern "C" int s2s_logmel(const float* wav, const float* window, const float* mel_basis, float* mel, int B,
                          int n_samples, int n_fft, int hop, int n_mels, float eps, float log_base, void* stream) {
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
    size_t smem = (size_t)half * 8 * 2 + (size_t)(nbins + 3) * 4 + (size_t)4 * nbins * 4 + (size_t)3 * n_mels * 4 + (size_t)n_fft * 4;
    static bool attr_set = false;
    if (smem > 48 * 1024 && !attr_set) {
        S2S_CUDA_OK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    S2S_REQUIRE(smem <= 160 * 1024, "logmel: shared memory request too large");
    long total = (long)B * n_frames;
    long grid = (long)num_sms() * 4;
    if (grid > total) grid = total;
    logmel_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(wav, window, mel_basis, mel, B, n_samples, n_fft, log2n - 1,
                                                                       hop, n_frames, n_mels, eps, log_scale);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
