// STFT -> log-mel front end (bin/preprocess.py:30-92 of the reference, librosa semantics).
//
// Fast path (n_fft = 2048): persistent CTA of 16 warps, one frame per warp and 16 frames per round.
//   The real 2048-point FFT is a 1024-point complex FFT of z[m] = x[2m] + i x[2m+1], computed as
//   32 x 32 Cooley-Tukey: every lane runs a 32-point FFT in registers over the stride-32 samples it
//   loaded straight from HBM (reflect padding and the window folded into the load), multiplies by the
//   inter-stage twiddles (a conflict-free 8 KB table), the warp transposes through its private 8 KB of
//   shared memory (__syncwarp only), every lane runs a second 32-point FFT, and the spectrum is unpacked
//   in conjugate pairs (one complex multiply per two bins) to |X[k]|, k = 0..1024.  The mel projection of the
//   CTA's 16 magnitude rows runs on warp-level tensor-core MMAs over the non-zero blocks of the banded
//   filterbank with bf16 high + low splits of both operands (fp32 accuracy), and 16 x n_mels logs leave as one
//   contiguous store.  Arithmetic: ~45 kFLOP per frame in fp32; each sample is re-used by n_fft / hop ~ 6.8
//   frames, so the kernel is bound by instruction issue / fp32 FFT arithmetic, not by its 1.52 kB / frame of HBM traffic.
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

// d * (c + i s) = (d.x c - d.y s, d.y c + d.x s)
__device__ __forceinline__ float2 cmul2(float2 d, float c, float s) {
    return fma2(make_float2(d.y, d.x), make_float2(-s, s), mul2(d, make_float2(c, c)));
}

// in-place N-point forward DFT (N = 8, 16, 32), decimation in frequency, fully unrolled; output v[brev(k)] = X[k] with brev the
// log2(N)-bit reversal
template <int N> __host__ __device__ constexpr int brevn(int x) {
    int r = 0;
    for (int b = 1, t = N >> 1; b < N; b <<= 1, t >>= 1)
        if (x & b) r |= t;
    return r;
}
template <int N>
__device__ __forceinline__ void fft_dif(float2 (&v)[32]) {
    constexpr int LOG = N == 32 ? 5 : (N == 16 ? 4 : 3);
#pragma unroll
    for (int s = 0; s < LOG; ++s) {
        const int half = (N / 2) >> s;         // butterfly span
#pragma unroll
        for (int g = 0; g < N; g += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const float2 a = v[g + j], b = v[g + j + half];
                v[g + j] = add2(a, b);
                const float2 d = sub2(a, b);
                const int tw = (j << s) * (32 / N);          // W_N^(j * 2^s) = W_32^tw
                if (tw == 0) v[g + j + half] = d;
                else if (tw == 8) v[g + j + half] = make_float2(d.y, -d.x);   // * (-i)
                else v[g + j + half] = cmul2(d, kCos32[tw], kSin32[tw]);
            }
        }
    }
}
__device__ __forceinline__ void fft32_dif(float2 (&v)[32]) { fft_dif<32>(v); }

// ---------------------------------------------------------------------------------------------
// n_fft = 2048 fast path: persistent CTA of 16 warps, 16 frames per round (one frame per warp)
// ---------------------------------------------------------------------------------------------
constexpr int LM_WARPS = 16;
constexpr int LM_ZLD = 33;                        // padded row length (float2) of the per-warp transpose / spectrum buffer
constexpr int LM_FS = 2120;                       // floats per warp region: >= 2 * 32 * 33, = 8 (mod 32) so the A-fragment LDS.64 are conflict-free
constexpr int LM_MAX_ITEMS = 96;                  // (8-band tile, 16-bin block) pairs of the mel projection
constexpr int LM_MAX_NT = 16;                     // 8-band tiles: n_mels <= 128 on the tensor-core mel path
constexpr int LM_MAX_PAIRS = LM_MAX_NT + LM_WARPS;
constexpr int LM_ITEMS_PER_WARP = LM_MAX_ITEMS / LM_WARPS;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x0, x1) -> bf16x2 high part (x0 in the low half) and bf16x2 of the remainders: x = hi + lo to ~2^-17 relative
__device__ __forceinline__ void split_bf16x2(float x0, float x1, unsigned& hi, unsigned& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct LogmelMeta {                               // built once per CTA from the filterbank matrix
    int n_items, use_tc;
    int item_nt[LM_MAX_ITEMS], item_k0[LM_MAX_ITEMS];
    int nt_lo[LM_MAX_NT], nt_cnt[LM_MAX_NT];
    int nt_np[LM_MAX_NT], nt_pair[LM_MAX_NT][LM_WARPS];     // partial tiles (offsets in floats) that make up a band tile, in warp order
    int chunk[LM_WARPS + 1], warp_pbase[LM_WARPS];
};

// Layout of dynamic shared memory (bytes): tw32 float2[1024] | tw2048 float2[520] | swin float[2048] (x 0.5) | btab uint4[96 * 32] |
// partial float[LM_MAX_PAIRS * 128] | meta | regions float[16 * LM_FS]
constexpr size_t LM_OFF_TW2048 = 1024 * 8;
constexpr size_t LM_OFF_WIN = LM_OFF_TW2048 + 520 * 8;
constexpr size_t LM_OFF_BTAB = LM_OFF_WIN + 2048 * 4;
constexpr size_t LM_OFF_PART = LM_OFF_BTAB + (size_t)LM_MAX_ITEMS * 32 * 16;
constexpr size_t LM_OFF_META = LM_OFF_PART + (size_t)LM_MAX_PAIRS * 128 * 4;
constexpr size_t LM_OFF_REG = (LM_OFF_META + sizeof(LogmelMeta) + 15) & ~(size_t)15;
constexpr size_t LM_SMEM = LM_OFF_REG + (size_t)LM_WARPS * LM_FS * 4;
static_assert(LM_SMEM <= 227 * 1024, "logmel2048: shared memory budget");

// Per round: every warp turns one frame into |X[0..1024]| (load + window -> 32x32 Cooley-Tukey in registers with one transpose
// through its own shared-memory region -> real-FFT unpack in conjugate pairs), the CTA then projects its 16 magnitude rows onto
// the mel bands with warp-level tensor-core MMAs (M = 16 frames, N = 8 bands, K = 16 bins; the filterbank is banded, so only
// the (band tile, bin block) pairs that hold non-zeros are visited; magnitudes and weights are split into bf16 high + low parts
// and three products are accumulated in fp32, which keeps the projection at fp32 accuracy), and 16 x n_mels logs are stored
// as one contiguous run.
template <int N1>      // n_fft = 64 * N1: 32 (2048), 16 (1024) or 8 (512); the complex FFT of H = 32 * N1 points is N1 x 32
__global__ void __launch_bounds__(LM_WARPS * 32, 1) logmel2048_kernel(const float* __restrict__ wav, const float* __restrict__ window,
                                                                     const float* __restrict__ basis, float* __restrict__ mel, int B,
                                                                     int ns, int hop, int n_frames, int n_mels, float eps,
                                                                     float log_scale, const float* __restrict__ nmean,
                                                                     const float* __restrict__ nscale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int H = 32 * N1, NB = H + 1;
    float2* tw32 = reinterpret_cast<float2*>(smem_raw);                       // [k1 * 32 + n2] = e^{-2 pi i k1 n2 / H}, k1 < N1
    float2* tw2048 = reinterpret_cast<float2*>(smem_raw + LM_OFF_TW2048);     // e^{-2 pi i k / (2 H)}, k <= H / 2
    float* swin = reinterpret_cast<float*>(smem_raw + LM_OFF_WIN);            // 0.5 * window: the 1/2 of the real-FFT unpack
    uint4* btab = reinterpret_cast<uint4*>(smem_raw + LM_OFF_BTAB);           // B fragments {b0 hi, b1 hi, b0 lo, b1 lo} per (item, lane)
    float* part = reinterpret_cast<float*>(smem_raw + LM_OFF_PART);           // per (warp, band tile) partial 16 x 8 tiles
    LogmelMeta& meta = *reinterpret_cast<LogmelMeta*>(smem_raw + LM_OFF_META);
    float* regions = reinterpret_cast<float*>(smem_raw + LM_OFF_REG);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    float2* zb = reinterpret_cast<float2*>(regions + (size_t)warp * LM_FS);
    float* mg = regions + (size_t)warp * LM_FS;

    // ---- CTA-wide tables
    for (int e = tid; e < 32 * N1; e += blockDim.x) {
        float s, c;
        sincospif(-2.0f * (float)((e >> 5) * (e & 31)) / (float)H, &s, &c);
        tw32[e] = make_float2(c, s);
    }
    for (int k = tid; k <= H / 2; k += blockDim.x) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)(2 * H), &s, &c);
        tw2048[k] = make_float2(c, s);
    }
    for (int k = tid; k < 2 * H; k += blockDim.x) swin[k] = 0.5f * window[k];
    const int NT = (n_mels + 7) >> 3;
    if (tid < LM_MAX_NT) { meta.nt_lo[tid] = NB; meta.nt_cnt[tid] = -1; }      // nt_cnt holds the last non-zero bin until the list is built
    __syncthreads();
    if (NT <= LM_MAX_NT) {
        // non-zero bin range of every 8-band tile: one warp per band row, all loads of a row in flight at once
        for (int m = warp; m < n_mels; m += LM_WARPS) {
            const float* row = basis + (size_t)m * NB;
            float w[33];
#pragma unroll
            for (int i = 0; i < 33; ++i) w[i] = (lane + 32 * i < NB) ? __ldg(row + lane + 32 * i) : 0.f;
            int lo = NB, hi = -1;
#pragma unroll
            for (int i = 0; i < 33; ++i)
                if (w[i] != 0.f) { lo = min(lo, lane + 32 * i); hi = max(hi, lane + 32 * i); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            if (lane == 0) { atomicMin(&meta.nt_lo[m >> 3], lo); atomicMax(&meta.nt_cnt[m >> 3], hi); }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        bool ok = NT <= LM_MAX_NT;
        for (int j = 0; ok && j < NT; ++j) {
            const int lo = meta.nt_lo[j] & ~1, hi = meta.nt_cnt[j];
            meta.nt_lo[j] = lo;
            meta.nt_cnt[j] = hi < 0 ? 0 : (hi + 1 - lo + 15) >> 4;
            if (n + meta.nt_cnt[j] > LM_MAX_ITEMS) { ok = false; break; }
            for (int c = 0; c < meta.nt_cnt[j]; ++c) { meta.item_nt[n] = j; meta.item_k0[n] = meta.nt_lo[j] + 16 * c; ++n; }
        }
        meta.use_tc = ok ? 1 : 0;
        meta.n_items = ok ? n : 0;
        if (ok) {
            for (int j = 0; j < NT; ++j) meta.nt_np[j] = 0;
            int pb = 0;
            for (int w = 0; w <= LM_WARPS; ++w) meta.chunk[w] = (w * n) / LM_WARPS;
            for (int w = 0; w < LM_WARPS; ++w) {
                const int i0 = meta.chunk[w], i1 = meta.chunk[w + 1];
                meta.warp_pbase[w] = pb;
                for (int i = i0; i < i1; ++i) {
                    const int j = meta.item_nt[i];
                    if (i == i0 || j != meta.item_nt[i - 1]) meta.nt_pair[j][meta.nt_np[j]++] = (pb + j - meta.item_nt[i0]) * 128;
                }
                if (i0 < i1) pb += meta.item_nt[i1 - 1] - meta.item_nt[i0] + 1;   // <= NT + 16 in total
            }
        }
    }
    __syncthreads();
    const bool use_tc = meta.use_tc != 0;
    for (int e = tid; e < meta.n_items * 32; e += blockDim.x) {
        const int it = e >> 5, ln = e & 31;
        const int m = meta.item_nt[it] * 8 + (ln >> 2), k = meta.item_k0[it] + 2 * (ln & 3);
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int kk = k + (q & 1) + 8 * (q >> 1);
            w[q] = (m < n_mels && kk < NB) ? basis[(size_t)m * NB + kk] : 0.f;
        }
        uint4 f;
        split_bf16x2(w[0], w[1], f.x, f.z);
        split_bf16x2(w[2], w[3], f.y, f.w);
        btab[e] = f;
    }
    __syncthreads();

    const long total = (long)B * n_frames;
    const long n_groups = (total + LM_WARPS - 1) / LM_WARPS;
    // lane n2 holds z[32 n1 + n2] = (x[2m], x[2m+1]), n1 = 0..31 (256 contiguous bytes per n1); frame f covers
    // [f * hop - n_fft / 2, ...) (center = True) with reflect padding (n_samples > n_fft / 2: one bounce)
    float2 v[32];
    auto fetch = [&](long frame) {
        if (frame >= total) return;
        const int b = (int)(frame / n_frames), f = (int)(frame - (long)b * n_frames);
        const float* x = wav + (size_t)b * ns;
        const int start = f * hop - H;
        if (start >= 0 && start + 2 * H <= ns && (reinterpret_cast<uintptr_t>(x + start) & 7) == 0) {
            const float2* p = reinterpret_cast<const float2*>(x + start) + lane;
#pragma unroll
            for (int n1 = 0; n1 < N1; ++n1) v[n1] = __ldg(p + 32 * n1);
        } else {
#pragma unroll
            for (int n1 = 0; n1 < N1; ++n1) {
                int s0 = start + 2 * (32 * n1 + lane), s1 = s0 + 1;
                s0 = s0 < 0 ? -s0 : (s0 >= ns ? 2 * (ns - 1) - s0 : s0);
                s1 = s1 < 0 ? -s1 : (s1 >= ns ? 2 * (ns - 1) - s1 : s1);
                v[n1] = make_float2(__ldg(x + s0), __ldg(x + s1));
            }
        }
    };
    // band sums in a fixed order over the contributing warps, log, (normalise,): warp w stores frame w's row of group gq, the 16
    // rows of a group are one contiguous run.  Half of the warps store right after the projection, the other half after their
    // next frame's FFT (any time before the next projection overwrites the partial tiles): the two halves then run half a phase
    // apart, so the arithmetic-bound and the shared-memory-bound stretches of different warps overlap.
    auto store_rows = [&](long gq) {
        const long fq = gq * LM_WARPS + warp;
        if (fq >= total) return;
        float* out = mel + (size_t)fq * n_mels;
        for (int m0 = lane; m0 < n_mels; m0 += 96) {
            float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int m = m0 + 32 * q;
                if (m < n_mels) {
                    const int nt = m >> 3, np = meta.nt_np[nt];
                    const float* pp = part + warp * 8 + (m & 7);
#pragma unroll 1
                    for (int c = 0; c < np; ++c) acc[q] += pp[meta.nt_pair[nt][c]];
                }
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int m = m0 + 32 * q;
                if (m < n_mels) {
                    float val = __log2f(fmaxf(eps, acc[q])) * log_scale;
                    if (nmean) val = (val - nmean[m]) / nscale[m];      // StandardScaler.transform (bin/normalize.py:193) fused
                    out[m] = val;
                }
            }
        }
    };
    // this warp's share of the (band tile, bin block) list is the same in every round
    const int i0 = meta.chunk[warp], n_it = meta.chunk[warp + 1] - i0;
    int it_k0[LM_ITEMS_PER_WARP], it_nt[LM_ITEMS_PER_WARP];
#pragma unroll
    for (int c = 0; c < LM_ITEMS_PER_WARP; ++c) {
        it_k0[c] = c < n_it ? meta.item_k0[i0 + c] : 0;
        it_nt[c] = c < n_it ? meta.item_nt[i0 + c] : 0;
    }
    float* part_w = part + (size_t)meta.warp_pbase[warp] * 128;
    const bool defer = ((warp >> 2) & 1) != 0;
    long pend = -1;
    fetch((long)blockIdx.x * LM_WARPS + warp);
    for (long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const long fr = g * LM_WARPS + warp;
        if (fr < total) {
            // ---- window (the raw samples were fetched during the previous round's projection / store stages)
            {
                const float2* sw2 = reinterpret_cast<const float2*>(swin) + lane;
#pragma unroll
                for (int n1 = 0; n1 < N1; ++n1) {
                    const float2 w = sw2[32 * n1];
                    v[n1] = make_float2(v[n1].x * w.x, v[n1].y * w.y);
                }
            }
            // ---- H-point complex FFT as N1 x 32: in-register DFTs around one transpose
            if constexpr (N1 == 32) {
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {                  // one copy of the 32-point code for both passes
                    fft32_dif(v);                                       // v[brev5(k)] = sum_n v[n] W_32^(n k)
                    if (pass == 0) {
#pragma unroll
                        for (int k1 = 0; k1 < 32; ++k1) zb[k1 * LM_ZLD + lane] = cmul(v[brev5(k1)], tw32[k1 * 32 + lane]);
                        __syncwarp();
#pragma unroll
                        for (int n2 = 0; n2 < 32; ++n2) v[n2] = zb[lane * LM_ZLD + n2];    // lane = k1 now
                        __syncwarp();
                    }
                }
            } else {
                // N1 < 32: the first pass is an N1-point DFT per lane, the second a 32-point DFT on lanes k1 < N1 (the other
                // lanes idle through it: a 1024-point frame costs ~60 % of a 2048-point one)
                fft_dif<N1>(v);
#pragma unroll
                for (int k1 = 0; k1 < N1; ++k1) zb[k1 * LM_ZLD + lane] = cmul(v[brevn<N1>(k1)], tw32[k1 * 32 + lane]);
                __syncwarp();
#pragma unroll
                for (int n2 = 0; n2 < 32; ++n2) v[n2] = lane < N1 ? zb[lane * LM_ZLD + n2] : make_float2(0.f, 0.f);
                __syncwarp();
                fft32_dif(v);
            }
            // ---- lane k1 now holds Z[k1 + 32 k2] in v[brev5(k2)].  Real-FFT unpack in conjugate pairs (k, 1024 - k): with
            //      e = Z[k] + conj Z[H-k], t = W_2048^k (Z[k] - conj Z[H-k]): X[k] = e - i t and X[H-k] = conj(e + i t).  Lane k1 takes
            //      k = k1 + 32 k2, k2 < 16; the partner Z[H-k] is lane (32 - k1)'s value k2' = 31 - k2 (one shuffle), for k1 = 0 the
            //      lane's own k2' = (32 - k2) & 31; k = 512 pairs with itself (lane 0).  The warp's region becomes the magnitude row.
            {
                const int src = (N1 - (lane & (N1 - 1))) & (N1 - 1);
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    const int k = (lane & (N1 - 1)) + N1 * k2;
                    const float2 zk = v[brev5(k2)], up = v[brev5(31 - k2)], own = v[brev5((32 - k2) & 31)];
                    float2 zcc = make_float2(__shfl_sync(0xffffffffu, up.x, src), -__shfl_sync(0xffffffffu, up.y, src));
                    if (lane == 0) zcc = make_float2(own.x, -own.y);            // conj Z[H-k]
                    const float2 e = add2(zk, zcc);
                    const float2 t = cmul(sub2(zk, zcc), tw2048[k]);
                    const float2 xa = add2(e, make_float2(t.y, -t.x)), xb = sub2(e, make_float2(t.y, -t.x));
                    if (N1 == 32 || lane < N1) {
                        mg[k] = sqrt_approx(xa.x * xa.x + xa.y * xa.y);
                        mg[H - k] = sqrt_approx(xb.x * xb.x + xb.y * xb.y);
                    }
                }
            }
            const float2 z512 = v[brev5(16)];
            const float m512 = 2.f * sqrt_approx(z512.x * z512.x + z512.y * z512.y);
            if (lane == 0) mg[H / 2] = m512;
            else mg[H + lane] = 0.f;                                    // bins H + 1 .. H + 31 are read (times zero weights) by the last bin blocks
            fetch(fr + (long)gridDim.x * LM_WARPS);                     // next round's samples: in flight across the projection and the store
            if (!use_tc) {
                // dense fallback for filterbanks the banded table cannot hold: one warp per frame, every band over every bin
                __syncwarp();
                float* out = mel + (size_t)fr * n_mels;
                for (int m = 0; m < n_mels; ++m) {
                    float acc = 0.f;
                    for (int k = lane; k < NB; k += 32) acc = fmaf(mg[k], basis[(size_t)m * NB + k], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) {
                        float o = log2f(fmaxf(eps, acc)) * log_scale;
                        if (nmean) o = (o - nmean[m]) / nscale[m];
                        out[m] = o;
                    }
                }
                __syncwarp();
            }
        }
        if (!use_tc) continue;
        if (pend >= 0) { store_rows(pend); pend = -1; }
        __syncthreads();
        // ---- mel projection of the 16 magnitude rows: this warp's share of the (band tile, bin block) list
        {
            float d[4] = {0.f, 0.f, 0.f, 0.f}, dx[4] = {0.f, 0.f, 0.f, 0.f};     // hi x hi and the two cross products: independent chains
            float* ps = part_w;
            const float* r0 = regions + (size_t)gid * LM_FS + 2 * tig;
            const float* r1 = r0 + 8 * LM_FS;
            int prev = it_nt[0];
#pragma unroll
            for (int c = 0; c < LM_ITEMS_PER_WARP; ++c) {
                if (c < n_it) {
                    const int nt = it_nt[c], k0 = it_k0[c];
                    if (nt != prev) {
                        *reinterpret_cast<float2*>(ps + gid * 8 + 2 * tig) = make_float2(d[0] + dx[0], d[1] + dx[1]);
                        *reinterpret_cast<float2*>(ps + (gid + 8) * 8 + 2 * tig) = make_float2(d[2] + dx[2], d[3] + dx[3]);
                        ps += (size_t)(nt - prev) * 128;
#pragma unroll
                        for (int q = 0; q < 4; ++q) d[q] = dx[q] = 0.f;
                        prev = nt;
                    }
                    const float2 x0 = *reinterpret_cast<const float2*>(r0 + k0), x1 = *reinterpret_cast<const float2*>(r1 + k0);
                    const float2 x2 = *reinterpret_cast<const float2*>(r0 + k0 + 8), x3 = *reinterpret_cast<const float2*>(r1 + k0 + 8);
                    unsigned ah[4], al[4];
                    split_bf16x2(x0.x, x0.y, ah[0], al[0]);
                    split_bf16x2(x1.x, x1.y, ah[1], al[1]);
                    split_bf16x2(x2.x, x2.y, ah[2], al[2]);
                    split_bf16x2(x3.x, x3.y, ah[3], al[3]);
                    const uint4 bw = btab[(i0 + c) * 32 + lane];
                    mma_bf16_16816(d, ah, bw.x, bw.y);
                    mma_bf16_16816(dx, al, bw.x, bw.y);
                    mma_bf16_16816(dx, ah, bw.z, bw.w);
                }
            }
            if (n_it > 0) {
                *reinterpret_cast<float2*>(ps + gid * 8 + 2 * tig) = make_float2(d[0] + dx[0], d[1] + dx[1]);
                *reinterpret_cast<float2*>(ps + (gid + 8) * 8 + 2 * tig) = make_float2(d[2] + dx[2], d[3] + dx[3]);
            }
        }
        __syncthreads();
        if (defer) pend = g; else store_rows(g);
    }
    if (pend >= 0) store_rows(pend);
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
    if (n_fft == 2048 || n_fft == 1024 || n_fft == 512) {
        // 16 frames per round per persistent CTA; filterbanks the banded table cannot hold take the kernel's dense fallback
        S2S_REQUIRE((long)n_samples + n_fft < (1L << 30), "logmel: clip too long for 32-bit sample indices");
        long grid = (long)num_sms();
        long need = ceil_div_l(total, LM_WARPS);
        if (grid > need) grid = need;
#define S2S_LM_LAUNCH(N1)                                                                                                      \
        do {                                                                                                                   \
            S2S_CUDA_OK(cudaFuncSetAttribute(logmel2048_kernel<N1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LM_SMEM)); \
            logmel2048_kernel<N1><<<(unsigned)grid, LM_WARPS * 32, LM_SMEM, (cudaStream_t)stream>>>(                           \
                wav, window, mel_basis, mel, B, n_samples, hop, n_frames, n_mels, eps, log_scale, nmean, nscale);             \
        } while (0)
        if (n_fft == 2048) S2S_LM_LAUNCH(32);
        else if (n_fft == 1024) S2S_LM_LAUNCH(16);
        else S2S_LM_LAUNCH(8);
#undef S2S_LM_LAUNCH
        S2S_LAUNCH_OK();
        return S2S_OK;
    }
    size_t smem = (size_t)half * 8 * 2 + (size_t)(nbins + 3) * 4 + (size_t)4 * nbins * 4 + (size_t)3 * n_mels * 4 + (size_t)n_fft * 4;
    if (smem > 48 * 1024)      // the attribute is per device: set it on every call that needs it (cheap)
        S2S_CUDA_OK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    S2S_REQUIRE(smem <= 160 * 1024, "logmel: shared memory request too large");
    long grid = (long)num_sms() * 4;
    if (grid > total) grid = total;
    logmel_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(wav, window, mel_basis, mel, B, n_samples, n_fft, log2n - 1,
                                                                       hop, n_frames, n_mels, eps, log_scale, nmean, nscale);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
