// Griffin-Lim phase reconstruction (seq2seq_vc/vocoder/griffin_lim.py:52-106 of the reference, which calls librosa.griffinlim:
// fast Griffin-Lim with momentum over librosa.stft / librosa.istft, center = True).  The iteration
//     inverse = istft(S * angles); rebuilt = stft(inverse); angles = rebuilt - c * previous; angles /= |angles| + tiny
// is four kernels per round, all device-resident:
//   gl_istft_frames_kernel  one CTA per frame: half-spectrum S[t,k] * angles[t,k] -> N real samples through an N/2-point complex
//                           inverse FFT in shared memory (even / odd packing), times the synthesis window
//   gl_overlap_add_kernel   y[n] = sum_t frame[t][n - t hop] / sum_t w^2[n - t hop], centre n_fft / 2 trimmed (a gather: no atomics)
//   gl_stft_kernel          one CTA per frame: zero / reflect padded samples * window -> N/2-point complex FFT -> X[0 .. N/2]
//   gl_update_kernel        the momentum step and the renormalisation to unit phases
// n_fft is a power of two in [64, 4096]; everything is fp32.
#include "common.cuh"

namespace s2s {

__device__ __forceinline__ float2 gl_cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// in-place radix-2 decimation-in-time FFT over `half` complex points that were stored bit-reversed; tw[k] = e^{-2 pi i k / (2 half)};
// inverse: conjugated twiddles (no 1 / half scaling here)
__device__ __forceinline__ void gl_fft_smem(float2* data, const float2* tw, int half, int log2h, bool inverse) {
    for (int s = 0; s < log2h; ++s) {
        const int hs = 1 << s;
        for (int t = threadIdx.x; t < (half >> 1); t += blockDim.x) {
            const int j = t & (hs - 1);
            const int i0 = ((t >> s) << (s + 1)) + j, i1 = i0 + hs;
            float2 w = tw[j << (log2h - s)];                 // e^{-2 pi i j / (2 hs)} on the n_fft-point table: index j * n_fft / (2 hs)
            if (inverse) w.y = -w.y;
            const float2 u = data[i0], v = gl_cmul(data[i1], w);
            data[i0] = make_float2(u.x + v.x, u.y + v.y);
            data[i1] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

// spec = mag * angles (T, half + 1) -> frames (T, n_fft): irfft of every frame times the window.  numpy's irfft ignores the imaginary
// parts of the DC and Nyquist bins; so does this kernel.
__global__ void __launch_bounds__(256) gl_istft_frames_kernel(const float* __restrict__ mag, const float2* __restrict__ angles,
                                                              const float* __restrict__ window, float* __restrict__ frames, int T,
                                                              int n_fft, int log2h) {
    extern __shared__ __align__(16) unsigned char gl_smem[];
    const int half = n_fft >> 1, nb = half + 1;
    float2* data = reinterpret_cast<float2*>(gl_smem);      // [half]
    float2* tw = data + half;                               // [half]   e^{-2 pi i k / n_fft}
    float2* xs = tw + half;                                 // [half + 1] the frame's spectrum
    for (int k = threadIdx.x; k < half; k += blockDim.x) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)n_fft, &s, &c);
        tw[k] = make_float2(c, s);
    }
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        for (int k = threadIdx.x; k < nb; k += blockDim.x) {
            const float m = mag[(size_t)t * nb + k];
            float2 a = angles[(size_t)t * nb + k];
            if (k == 0 || k == half) a.y = 0.f;
            xs[k] = make_float2(m * a.x, m * a.y);
        }
        __syncthreads();
        // Z[k] = E[k] + i O[k],  E = (X[k] + conj X[half-k]) / 2,  O = conj(W^k) (X[k] - conj X[half-k]) / 2;  stored bit-reversed
        for (int k = threadIdx.x; k < half; k += blockDim.x) {
            const float2 a = xs[k];
            float2 b = xs[half - k];
            b.y = -b.y;
            const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
            float2 w = tw[k];
            w.y = -w.y;
            const float2 o = gl_cmul(make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y)), w);
            const unsigned r = __brev((unsigned)k) >> (32 - log2h);
            data[r] = make_float2(e.x - o.y, e.y + o.x);
        }
        __syncthreads();
        gl_fft_smem(data, tw, half, log2h, true);
        const float sc = 1.f / (float)half;
        float* out = frames + (size_t)t * n_fft;
        for (int m = threadIdx.x; m < half; m += blockDim.x) {
            const float2 z = data[m];
            out[2 * m] = z.x * sc * window[2 * m];
            out[2 * m + 1] = z.y * sc * window[2 * m + 1];
        }
        __syncthreads();
    }
}

// y[j] = (sum_t frames[t][n - t hop]) / (sum_t window[n - t hop]^2),  n = j + n_fft / 2,  j < hop (T - 1)
__global__ void __launch_bounds__(256) gl_overlap_add_kernel(const float* __restrict__ frames, const float* __restrict__ window,
                                                             float* __restrict__ y, int T, int n_fft, int hop, long n_out, float tiny) {
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (long)gridDim.x * blockDim.x) {
        const long n = j + (n_fft >> 1);
        long t_hi = n / hop;
        if (t_hi > T - 1) t_hi = T - 1;
        long t_lo = 0;                                       // smallest t with n - t hop < n_fft
        if (n - n_fft + 1 > 0) t_lo = (n - n_fft + 1 + hop - 1) / hop;
        float acc = 0.f, wss = 0.f;
        for (long t = t_lo; t <= t_hi; ++t) {
            const int q = (int)(n - t * hop);
            const float w = window[q];
            acc += frames[(size_t)t * n_fft + q];
            wss = fmaf(w, w, wss);
        }
        y[j] = wss > tiny ? acc / wss : acc;
    }
}

// X[t, k], k = 0 .. half, of the centred frames of y (n_samples); pad_reflect = 0: zero padding ("constant"), 1: reflect
__global__ void __launch_bounds__(256) gl_stft_kernel(const float* __restrict__ y, const float* __restrict__ window,
                                                      float2* __restrict__ spec, int T, long ns, int n_fft, int log2h, int hop,
                                                      int pad_reflect) {
    extern __shared__ __align__(16) unsigned char gl_smem[];
    const int half = n_fft >> 1, nb = half + 1;
    float2* data = reinterpret_cast<float2*>(gl_smem);      // [half]
    float2* tw = data + half;                               // [half]
    for (int k = threadIdx.x; k < half; k += blockDim.x) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)n_fft, &s, &c);
        tw[k] = make_float2(c, s);
    }
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        const long start = (long)t * hop - half;
        for (int m = threadIdx.x; m < half; m += blockDim.x) {
            float v[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                long s = start + 2 * m + q;
                float x = 0.f;
                if (pad_reflect) {
                    if (s < 0) s = -s; else if (s >= ns) s = 2 * (ns - 1) - s;
                    x = (s >= 0 && s < ns) ? y[s] : 0.f;
                } else if (s >= 0 && s < ns) {
                    x = y[s];
                }
                v[q] = x * window[2 * m + q];
            }
            const unsigned r = __brev((unsigned)m) >> (32 - log2h);
            data[r] = make_float2(v[0], v[1]);
        }
        __syncthreads();
        gl_fft_smem(data, tw, half, log2h, false);
        float2* out = spec + (size_t)t * nb;
        for (int k = threadIdx.x; k <= half; k += blockDim.x) {
            const float2 zk = data[k & (half - 1)];
            float2 zc = data[(half - k) & (half - 1)];
            zc.y = -zc.y;
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            const float2 o = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
            const float2 w = (k < half) ? tw[k] : make_float2(-1.f, 0.f);
            const float2 ow = gl_cmul(o, w);                // X[k] = E[k] + W^k O[k],  O = -i o
            out[k] = make_float2(e.x + ow.y, e.y - ow.x);
        }
        __syncthreads();
    }
}

// angles = rebuilt - c * tprev;  angles /= |angles| + tiny;  tprev = rebuilt
__global__ void __launch_bounds__(256) gl_update_kernel(const float2* __restrict__ rebuilt, float2* __restrict__ tprev,
                                                        float2* __restrict__ angles, long n, float c, float tiny) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float2 r = rebuilt[i], p = tprev[i];
        const float2 a = make_float2(r.x - c * p.x, r.y - c * p.y);
        const float inv = 1.f / (sqrtf(a.x * a.x + a.y * a.y) + tiny);
        angles[i] = make_float2(a.x * inv, a.y * inv);
        tprev[i] = r;
    }
}

}  // namespace s2s

using namespace s2s;

static int gl_log2h(int n_fft) {
    int l = 0;
    while ((1 << l) < n_fft) ++l;
    return l - 1;
}

extern "C" int s2s_gl_istft(const float* mag, const float* angles, const float* window, float* frames, float* y, int T, int n_fft,
                            int hop, void* stream) {
    S2S_REQUIRE(mag && angles && window && frames && y && T >= 2 && hop > 0, "gl_istft: bad arguments");
    S2S_REQUIRE(n_fft >= 64 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "gl_istft: n_fft %d must be a power of two in [64, 4096]", n_fft);
    cudaStream_t st = (cudaStream_t)stream;
    const int half = n_fft / 2;
    const size_t smem = ((size_t)2 * half + half + 1) * sizeof(float2);
    if (smem > 48 * 1024) S2S_CUDA_OK(cudaFuncSetAttribute(gl_istft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    int grid = T < num_sms() * 4 ? T : num_sms() * 4;
    gl_istft_frames_kernel<<<grid, 256, smem, st>>>(mag, reinterpret_cast<const float2*>(angles), window, frames, T, n_fft, gl_log2h(n_fft));
    S2S_LAUNCH_OK();
    const long n_out = (long)hop * (T - 1);
    gl_overlap_add_kernel<<<(unsigned)ceil_div_l(n_out, 256), 256, 0, st>>>(frames, window, y, T, n_fft, hop, n_out, 1.17549435e-38f);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_gl_stft(const float* y, const float* window, float* spec, int T, int64_t n_samples, int n_fft, int hop,
                           int pad_reflect, void* stream) {
    S2S_REQUIRE(y && window && spec && T >= 1 && hop > 0 && n_samples > 0, "gl_stft: bad arguments");
    S2S_REQUIRE(n_fft >= 64 && n_fft <= 4096 && (n_fft & (n_fft - 1)) == 0, "gl_stft: n_fft %d must be a power of two in [64, 4096]", n_fft);
    S2S_REQUIRE(!pad_reflect || n_samples > n_fft / 2, "gl_stft: reflect padding needs n_samples > n_fft / 2");
    const int half = n_fft / 2;
    const size_t smem = (size_t)2 * half * sizeof(float2);
    int grid = T < num_sms() * 4 ? T : num_sms() * 4;
    gl_stft_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(y, window, reinterpret_cast<float2*>(spec), T, (long)n_samples, n_fft,
                                                              gl_log2h(n_fft), hop, pad_reflect);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_gl_update(const float* rebuilt, float* tprev, float* angles, int64_t n, float c, void* stream) {
    S2S_REQUIRE(rebuilt && tprev && angles && n > 0, "gl_update: bad arguments");
    long grid = ceil_div_l(n, 256);
    if (grid > (long)num_sms() * 8) grid = (long)num_sms() * 8;
    gl_update_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(rebuilt), reinterpret_cast<float2*>(tprev),
                                                                      reinterpret_cast<float2*>(angles), n, c, 1.17549435e-38f);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
