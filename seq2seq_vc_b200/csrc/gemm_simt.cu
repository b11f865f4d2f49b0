// fp32 CUDA-core GEMM: the parity ("exact") path behind s2s_gemm(mode=0).
//
// 128x128x16 block tile, 256 threads, 8x8 register micro-tile split as 2x2 groups of 4x4 so the
// float4 shared-memory reads are bank-conflict free.  Operands may be f32 or bf16 in HBM with any
// (row, col) element strides (one of them 1); accumulation is fp32 FMA in k order inside a tile.
// This kernel is deliberately simple: it is the numerical yard-stick the tcgen05 path is compared
// against on the GPU, and the fp32 mode used for the reference-parity configuration (C1).
#include "common.cuh"

namespace s2s {

constexpr int SBM = 128, SBN = 128, SBK = 16, SPAD = 4;

struct SimtParams {
    s2s_gemm_t g;
    Dropout drop;
    long KK;  // taps * K
};

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
    __shared__ __align__(16) float As[SBK][SBM + SPAD];
    __shared__ __align__(16) float Bs[SBK][SBN + SPAD];
    const s2s_gemm_t& g = p.g;
    Dropout drop = p.drop;
    dropout_resolve(drop);
    const int tid = threadIdx.x;
    const int bz = blockIdx.z;
    const int b1 = bz / g.batch2, b2 = bz % g.batch2;
    const TA* __restrict__ A = reinterpret_cast<const TA*>(g.A) + b1 * g.a_bs1 + b2 * g.a_bs2;
    const TB* __restrict__ B = reinterpret_cast<const TB*>(g.B) + b1 * g.b_bs1 + b2 * g.b_bs2;
    TC* __restrict__ C = reinterpret_cast<TC*>(g.C) + b1 * g.c_bs1 + b2 * g.c_bs2;
    const TC* __restrict__ R = g.R ? reinterpret_cast<const TC*>(g.R) + b1 * g.c_bs1 + b2 * g.c_bs2 : nullptr;
    const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;
    const int tx = tid & 15, ty = tid >> 4;
    const bool a_kmajor = (g.a_cs == 1), b_kmajor = (g.b_cs == 1);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (long kk0 = 0; kk0 < p.KK; kk0 += SBK) {
        // ---- stage A tile (128 x 16) and B tile (128 x 16) into shared memory, k-major
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int idx = tid + i * 256;
            int k, r;
            if (a_kmajor) { k = idx & 15; r = idx >> 4; } else { r = idx & 127; k = idx >> 7; }
            long kk = kk0 + k;
            int m = m0 + r;
            float v = 0.f;
            if (kk < p.KK && m < g.M) {
                int t = (int)(kk / g.K);
                int kq = (int)(kk - (long)t * g.K);
                v = to_f<TA>(A[(long)(m + t) * g.a_rs + (long)kq * g.a_cs]);
            }
            As[k][r] = v;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int idx = tid + i * 256;
            int k, r;
            if (b_kmajor) { k = idx & 15; r = idx >> 4; } else { r = idx & 127; k = idx >> 7; }
            long kk = kk0 + k;
            int n = n0 + r;
            float v = 0.f;
            if (kk < p.KK && n < g.N) {
                int t = (int)(kk / g.K);
                int kq = (int)(kk - (long)t * g.K);
                v = to_f<TB>(B[(long)n * g.b_rs + (long)t * g.b_ts + (long)kq * g.b_cs]);
            }
            Bs[k][r] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SBK; ++k) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue
    const long batch_lin = (long)bz * g.M;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
        if (m >= g.M) continue;
        bool row_ok = true;
        if (g.mask_period > 0) {
            int ph = (m + g.mask_offset) % g.mask_period;
            row_ok = (ph >= g.mask_lo) && (ph < g.mask_hi);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int n = n0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4)));
            if (n >= g.N) continue;
            float v = acc[i][j] * g.alpha;
            if (g.bias) v += g.bias[n];
            if (g.relu) v = fmaxf(v, 0.f);
            v *= dropout_factor(drop, (uint64_t)((batch_lin + m) * (long)g.N + n));
            TC* dst = C + (long)m * g.c_rs + n;
            if (R) {
                const float r = to_f<TC>(R[(long)m * g.c_rs + n]);
                v = g.r_mode ? (r > 0.f ? v * g.r_scale : 0.f) : v + r;
            }
            if (g.accumulate) v += to_f<TC>(*dst);
            if (!row_ok) v = 0.f;
            *dst = from_f<TC>(v);
        }
    }
}

template <typename TA, typename TB, typename TC>
static int launch_simt(const s2s_gemm_t& g, cudaStream_t st) {
    SimtParams p;
    p.g = g;
    p.drop = make_dropout(&g.drop);
    p.KK = (long)g.taps * g.K;
    dim3 grid((unsigned)ceil_div_l(g.M, SBM), (unsigned)ceil_div_l(g.N, SBN), (unsigned)(g.batch1 * g.batch2));
    gemm_simt_kernel<TA, TB, TC><<<grid, 256, 0, st>>>(p);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

int gemm_simt(const s2s_gemm_t& g, cudaStream_t st) {
#define S2S_SIMT_CASE(da, db, dc, TA, TB, TC) \
    if (g.a_dtype == da && g.b_dtype == db && g.c_dtype == dc) return launch_simt<TA, TB, TC>(g, st);
    S2S_SIMT_CASE(S2S_F32, S2S_F32, S2S_F32, float, float, float)
    S2S_SIMT_CASE(S2S_BF16, S2S_BF16, S2S_F32, bf16, bf16, float)
    S2S_SIMT_CASE(S2S_BF16, S2S_BF16, S2S_BF16, bf16, bf16, bf16)
    S2S_SIMT_CASE(S2S_F32, S2S_F32, S2S_BF16, float, float, bf16)
#undef S2S_SIMT_CASE
    return set_error(S2S_ERR_UNSUPPORTED, "gemm_simt: unsupported dtype combination a=%d b=%d c=%d", g.a_dtype,
                     g.b_dtype, g.c_dtype);
}

}  // namespace s2s
