// Flash-style fused multi-head attention on the 5th-generation tensor cores (tcgen05 + TMEM), forward and backward.
//
//   reference: MultiHeadedAttention.forward_attention (modules/transformer/attention.py:76-111):
//       scores = q k^T / sqrt(d_k); masked_fill(mask == 0, min); softmax; masked_fill(mask == 0, 0.0); x = attn v
//   with the padding mask given as a per-utterance key length and the decoder's subsequent mask as a causal flag
//   (models/vtn.py:553-602).  Masked columns are exactly zero after the softmax; a row without any visible key gives
//   an all-zero probability row and a zero context row.
//
// The (B,H,T1,T2) score / probability matrix never touches HBM (unless the caller asks for it: the source-attention
// maps are an output of VTN.forward, models/vtn.py:280-287): S lives in tensor memory, P goes TMEM -> registers ->
// shared memory (bf16, the K-major SWIZZLE_128B layout tcgen05.mma reads) -> second MMA.  The backward recomputes P
// from Q, K and the saved row statistics.
//
// One CTA = one 128-row tile of one (batch, head); the other sequence is streamed in blocks of 64 rows through a
// 2-stage TMA ring.  192 threads:
//   warps 0-3  compute: thread t owns row t of the tile = TMEM lane t (row statistics need no shuffles)
//   warp 4     TMA producer (cp.async.bulk.tensor 4-D boxes of 64 bf16 x rows, 128B swizzle; d_k tails zero-filled)
//   warp 5     MMA issuer (one elected lane), owns the TMEM allocation
// Three kernels share this skeleton (template MODE); per streamed block
//   first-stage MMAs  X1 = R1 C1^T (, X2 = R2 C2^T)            128 x 64 x d_k  -> TMEM
//   compute threads   X -> bf16 operand tiles W1 (, W2) in shared memory
//   second-stage MMAs acc += W C                                128 x d_k x 64  -> TMEM accumulators
//   MODE_FWD    R1 = Q tile, C1 = K block, C2 = V block: pass 1 row max (and sum when P is emitted), pass 2
//               P = exp2(S - m) -> W1, O += P V; epilogue O / l -> ctx, lse = m + log2 l
//   MODE_BWD_Q  R1 = Q, R2 = dO, C1 = K, C2 = V: S, dP = dO V^T; dS = P (dP - D) scale -> W1; dQ += dS K
//   MODE_BWD_KV R1 = K tile, R2 = V tile, C1 = Q block, C2 = dO block: S^T = K Q^T, dP^T = V dO^T; P^T -> W1,
//               dS^T -> W2; dV += P^T dO, dK += dS^T Q
// Without look-ahead inside a CTA (X is single-buffered): two or three CTAs share an SM and overlap each other's
// MMA / exp phases; the tensor work of a d_k = 48 head is a few percent of the exp / convert work anyway.
#include "tc_common.cuh"

namespace s2s {
namespace atc {
using namespace tc;

constexpr int BM = 128, BN = 64, NST = 2, NUM_THREADS = 192;
constexpr int MODE_FWD = 0, MODE_BWD_Q = 1, MODE_BWD_KV = 2;
constexpr int ROW_PANEL = BM * 128;      // bytes of one 64-column panel of a 128-row operand tile
constexpr int BLK_PANEL = BN * 128;      // bytes of one 64-column panel of a 64-row streamed block

struct View {                // element (b, t, h, j) at p + b * bs + t * ts + h * hs + j
    bf16* p;
    long bs, ts, hs;
};

struct Params {
    CUtensorMap tmR1, tmR2, tmC1, tmC2;
    int B, H, T1, T2, dk, KA;       // KA = 64-column panels along d_k (1 for d_k <= 64, 2 up to 128)
    int T1p;                        // row pitch of lse / D (T1 rounded up to 64)
    int causal;
    float scale, scale_log2;        // 1/sqrt(d_k) and the same times log2(e)
    const int32_t* klens;
    View ctx;                       // fwd: out; bwd_q: O (for D = rowsum(dO o O))
    View dO;                        // bwd_q
    View out1, out2;                // bwd_q: dQ; bwd_kv: dV, dK
    float* lse;                     // (B,H,T1p) log2-domain row statistics m + log2(l); +inf for rows without a visible key
    float* Dvec;                    // (B,H,T1p) rowsum(dO o O): written by bwd_q, read by bwd_kv
    bf16* P;                        // fwd, optional: normalised probabilities (B,H,T1,ld)
    long ld;
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// 32 values (columns c0 .. c0 + 31 of this thread's row) -> four 16-byte chunks of the K-major SWIZZLE_128B operand tile
__device__ __forceinline__ void store_w32(uint32_t wbase, int row, int c0, const float (&v)[32]) {
    const uint32_t rowaddr = wbase + (uint32_t)row * 128u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint4 w;
        w.x = pack2(v[8 * g + 0], v[8 * g + 1]);
        w.y = pack2(v[8 * g + 2], v[8 * g + 3]);
        w.z = pack2(v[8 * g + 4], v[8 * g + 5]);
        w.w = pack2(v[8 * g + 6], v[8 * g + 7]);
        const int chunk = (c0 >> 3) + g;
        st_shared_v4(rowaddr + (uint32_t)((chunk ^ (row & 7)) << 4), w);
    }
}

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 2) attn_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    constexpr bool BWD = MODE != MODE_FWD;
    const int KA = p.KA;
    // shared-memory map
    const uint32_t sR1 = base;
    const uint32_t sR2 = sR1 + (uint32_t)KA * ROW_PANEL;
    const uint32_t sST = sR2 + (BWD ? (uint32_t)KA * ROW_PANEL : 0u);
    const uint32_t stage_bytes = 2u * (uint32_t)KA * BLK_PANEL;               // C1 panels then C2 panels
    const uint32_t sW1 = sST + NST * stage_bytes;
    const uint32_t sW2 = sW1 + BM * 128;
    const uint32_t bars = sW2 + (MODE == MODE_BWD_KV ? BM * 128 : 0);
    const uint32_t row_full = bars, x_full = bars + 8, w_full = bars + 16, acc_full = bars + 24;
    auto st_full = [&](int s) { return bars + 32u + 8u * s; };
    auto st_empty = [&](int s) { return bars + 32u + 8u * (NST + s); };
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y;
    const int r0 = blockIdx.x * BM;                         // first row of this CTA's tile (query rows, or key rows for BWD_KV)
    int klen = p.klens ? p.klens[b] : p.T2;
    klen = klen < 0 ? 0 : (klen > p.T2 ? p.T2 : klen);

    // streamed blocks this CTA visits: [blk0, blk1)
    int blk0 = 0, blk1 = 0;
    if (MODE == MODE_BWD_KV) {
        if (r0 < klen) {
            blk0 = p.causal ? r0 / BN : 0;
            blk1 = (p.T1 + BN - 1) / BN;
        }
    } else {
        const int last_row = min(p.T1, r0 + BM) - 1;
        const int lim = p.causal ? min(klen, last_row + 1) : klen;
        blk1 = (lim + BN - 1) / BN;
    }
    const int nblk = blk1 - blk0;
    const int n_iter = (MODE == MODE_FWD) ? 2 * nblk : nblk;
    const int dk = p.dk, DKP = 64 * KA;
    // tensor-memory columns: X1 [0,64) | X2 [64,128) (bwd) | acc1 | acc2
    const uint32_t colX2 = 64, colA1 = BWD ? 128 : 64, colA2 = colA1 + (uint32_t)DKP;
    uint32_t tmem_cols = colA1 + (uint32_t)DKP * (MODE == MODE_BWD_KV ? 2 : 1);
    tmem_cols = tmem_cols <= 128 ? 128 : (tmem_cols <= 256 ? 256 : 512);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmR1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmC1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmC2)) : "memory");
        mbar_init(row_full, 1);
        mbar_init(x_full, 1);
        mbar_init(w_full, 4);
        mbar_init(acc_full, 1);
        for (int s = 0; s < NST; ++s) { mbar_init(st_full(s), 1); mbar_init(st_empty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 4) {
        // ================= TMA producer =================
        if (n_iter > 0 && elect_one()) {
            const uint32_t row_bytes = (uint32_t)KA * ROW_PANEL * (BWD ? 2u : 1u);
            mbar_expect_tx(row_full, row_bytes);
            for (int a = 0; a < KA; ++a) {
                tma_load_4d(sR1 + a * ROW_PANEL, &p.tmR1, row_full, a * 64, r0, h, b);
                if (BWD) tma_load_4d(sR2 + a * ROW_PANEL, &p.tmR2, row_full, a * 64, r0, h, b);
            }
            int s = 0;
            uint32_t ph = 0;
#pragma unroll 1
            for (int it = 0; it < n_iter; ++it) {
                const bool pass1 = (MODE == MODE_FWD) && it < nblk;
                const int blk = blk0 + ((MODE == MODE_FWD && it >= nblk) ? it - nblk : it);
                mbar_wait(st_empty(s), ph ^ 1u);
                const uint32_t sC1 = sST + (uint32_t)s * stage_bytes, sC2 = sC1 + (uint32_t)KA * BLK_PANEL;
                mbar_expect_tx(st_full(s), (uint32_t)KA * BLK_PANEL * (pass1 ? 1u : 2u));
                for (int a = 0; a < KA; ++a) {
                    tma_load_4d(sC1 + a * BLK_PANEL, &p.tmC1, st_full(s), a * 64, blk * BN, h, b);
                    if (!pass1) tma_load_4d(sC2 + a * BLK_PANEL, &p.tmC2, st_full(s), a * 64, blk * BN, h, b);
                }
                if (++s == NST) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (n_iter > 0) {
            // first stage: M = 128, N = 64, both operands K-major; second stage: M = 128, N = d_k, A K-major, B MN-major
            const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(dk >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const int ksteps1 = dk / 16;
            mbar_wait(row_full, 0);
            int s = 0;
            uint32_t ph = 0;
            bool acc_started = false;
#pragma unroll 1
            for (int it = 0; it < n_iter; ++it) {
                const bool pass1 = (MODE == MODE_FWD) && it < nblk;
                const bool last = it == n_iter - 1;
                const uint32_t sC1 = sST + (uint32_t)s * stage_bytes, sC2 = sC1 + (uint32_t)KA * BLK_PANEL;
                mbar_wait(st_full(s), ph);
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll 1
                    for (int ks = 0; ks < ksteps1; ++ks) {
                        const uint32_t ao = (uint32_t)(ks >> 2) * ROW_PANEL + (uint32_t)(ks & 3) * 32u;
                        const uint32_t bo = (uint32_t)(ks >> 2) * BLK_PANEL + (uint32_t)(ks & 3) * 32u;
                        umma_bf16(tmem_base, smem_desc(sR1 + ao, 16, 1024), smem_desc(sC1 + bo, 16, 1024), idesc1, ks > 0 ? 1u : 0u);
                        if (BWD) umma_bf16(tmem_base + colX2, smem_desc(sR2 + ao, 16, 1024), smem_desc(sC2 + bo, 16, 1024), idesc1, ks > 0 ? 1u : 0u);
                    }
                    umma_commit(x_full);
                    if (pass1) umma_commit(st_empty(s));
                }
                __syncwarp();
                mbar_wait(w_full, (uint32_t)(it & 1));
                tcgen05_fence_after();
                if (!pass1) {
                    if (elect_one()) {
                        const uint32_t accf = acc_started ? 1u : 0u;
#pragma unroll 1
                        for (int ks = 0; ks < BN / 16; ++ks) {
                            const uint64_t wd1 = smem_desc(sW1 + (uint32_t)ks * 32u, 16, 1024);
                            const uint32_t co = (uint32_t)ks * 2048u;
                            if (MODE == MODE_FWD) {
                                umma_bf16(tmem_base + colA1, wd1, smem_desc(sC2 + co, BLK_PANEL, 1024), idesc2, (accf | (ks > 0)) ? 1u : 0u);
                            } else if (MODE == MODE_BWD_Q) {
                                umma_bf16(tmem_base + colA1, wd1, smem_desc(sC1 + co, BLK_PANEL, 1024), idesc2, (accf | (ks > 0)) ? 1u : 0u);
                            } else {
                                umma_bf16(tmem_base + colA1, wd1, smem_desc(sC2 + co, BLK_PANEL, 1024), idesc2, (accf | (ks > 0)) ? 1u : 0u);
                                umma_bf16(tmem_base + colA2, smem_desc(sW2 + (uint32_t)ks * 32u, 16, 1024), smem_desc(sC1 + co, BLK_PANEL, 1024), idesc2,
                                          (accf | (ks > 0)) ? 1u : 0u);
                            }
                        }
                        umma_commit(st_empty(s));
                        if (last) umma_commit(acc_full);
                    }
                    __syncwarp();
                    acc_started = true;
                }
                if (++s == NST) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ================= compute threads: thread t <-> row r0 + t <-> TMEM lane t =================
        const int t = threadIdx.x;                       // 0..127
        const int row = r0 + t;
        const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
        const long bh = (long)b * p.H + h;
        auto signal_w = [&]() {                          // this warp is done with X (and its slice of W is visible to the async proxy)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(w_full);
        };

        if (MODE == MODE_FWD) {
            const bool row_in = row < p.T1;
            const int lim_row = p.causal ? min(klen, row + 1) : klen;       // visible keys of this row: [0, lim_row)
            const bool emit = p.P != nullptr;
            bf16* Prow = emit ? p.P + (bh * p.T1 + row) * p.ld : nullptr;
            float m = -INFINITY, l = 0.f;
            int it = 0;
            // ---- pass 1: row maximum (and, when P is emitted, the row sum: the probabilities leave normalised)
#pragma unroll 1
            for (int j = 0; j < nblk; ++j, ++it) {
                mbar_wait(x_full, (uint32_t)(it & 1));
                tcgen05_fence_after();
                uint32_t a0[32], a1[32];
                tmem_ld32(tlane, a0);
                tmem_ld32(tlane + 32, a1);
                tmem_ld_wait();
                signal_w();
                const int rem = lim_row - j * BN;        // columns idx < rem are visible
                float bm = -INFINITY;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float v0 = (c < rem) ? __uint_as_float(a0[c]) * p.scale_log2 : -INFINITY;
                    const float v1 = (c + 32 < rem) ? __uint_as_float(a1[c]) * p.scale_log2 : -INFINITY;
                    a0[c] = __float_as_uint(v0);
                    a1[c] = __float_as_uint(v1);
                    bm = fmaxf(bm, fmaxf(v0, v1));
                }
                if (emit && bm > -INFINITY) {
                    const float mn = fmaxf(m, bm);
                    float sum = 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) sum += ex2f(__uint_as_float(a0[c]) - mn) + ex2f(__uint_as_float(a1[c]) - mn);
                    l = l * ex2f(m - mn) + sum;
                    m = mn;
                } else {
                    m = fmaxf(m, bm);
                }
            }
            const float inv_l1 = (emit && l > 0.f) ? 1.f / l : 0.f;
            float l2 = 0.f;
            // ---- pass 2: probabilities -> W1 (-> P), O += P V
#pragma unroll 1
            for (int j = 0; j < nblk; ++j, ++it) {
                mbar_wait(x_full, (uint32_t)(it & 1));
                tcgen05_fence_after();
                const int rem = lim_row - j * BN;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t a[32];
                    tmem_ld32(tlane + half * 32, a);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        float e = ex2f(fmaf(__uint_as_float(a[c]), p.scale_log2, -m));
                        e = (c + half * 32 < rem) ? e : 0.f;
                        l2 += e;
                        v[c] = emit ? e * inv_l1 : e;
                    }
                    store_w32(sW1, t, half * 32, v);
                    if (emit && row_in) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int col = j * BN + half * 32 + g * 8;
                            if (col < p.ld) {
                                uint4 w;
                                w.x = pack2(v[8 * g], v[8 * g + 1]); w.y = pack2(v[8 * g + 2], v[8 * g + 3]);
                                w.z = pack2(v[8 * g + 4], v[8 * g + 5]); w.w = pack2(v[8 * g + 6], v[8 * g + 7]);
                                *reinterpret_cast<uint4*>(Prow + col) = w;
                            }
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                signal_w();
            }
            if (emit && row_in) {       // key blocks no row of this tile can see, up to the padded row length: zeros
                for (int col = nblk * BN + 0; col < p.ld; col += 8) *reinterpret_cast<uint4*>(Prow + col) = make_uint4(0u, 0u, 0u, 0u);
            }
            // ---- epilogue: O (/ l) -> ctx, row statistics for the backward
            const float lsum = emit ? l : l2;
            const float inv = emit ? 1.f : (lsum > 0.f ? 1.f / lsum : 0.f);
            if (row_in) p.lse[bh * p.T1p + row] = (nblk > 0 && lsum > 0.f) ? m + log2f(lsum) : INFINITY;
            bf16* orow = p.ctx.p + (long)b * p.ctx.bs + (long)row * p.ctx.ts + (long)h * p.ctx.hs;
            if (nblk > 0) {
                mbar_wait(acc_full, 0);
                tcgen05_fence_after();
            }
#pragma unroll 1
            for (int c0 = 0; c0 < dk; c0 += 16) {
                uint32_t a[16];
                if (nblk > 0) {
                    tmem_ld16(tlane + 64 + c0, a);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) a[c] = 0u;
                }
                if (row_in) {
                    uint4 w0, w1;
                    w0.x = pack2(__uint_as_float(a[0]) * inv, __uint_as_float(a[1]) * inv); w0.y = pack2(__uint_as_float(a[2]) * inv, __uint_as_float(a[3]) * inv);
                    w0.z = pack2(__uint_as_float(a[4]) * inv, __uint_as_float(a[5]) * inv); w0.w = pack2(__uint_as_float(a[6]) * inv, __uint_as_float(a[7]) * inv);
                    w1.x = pack2(__uint_as_float(a[8]) * inv, __uint_as_float(a[9]) * inv); w1.y = pack2(__uint_as_float(a[10]) * inv, __uint_as_float(a[11]) * inv);
                    w1.z = pack2(__uint_as_float(a[12]) * inv, __uint_as_float(a[13]) * inv); w1.w = pack2(__uint_as_float(a[14]) * inv, __uint_as_float(a[15]) * inv);
                    *reinterpret_cast<uint4*>(orow + c0) = w0;
                    *reinterpret_cast<uint4*>(orow + c0 + 8) = w1;
                }
            }
        } else if (MODE == MODE_BWD_Q) {
            const bool row_in = row < p.T1;
            const int lim_row = p.causal ? min(klen, row + 1) : klen;
            // D = rowsum(dO o O) and the saved row statistic
            float D = 0.f, lse = INFINITY;
            if (row_in) {
                const bf16* orow = p.ctx.p + (long)b * p.ctx.bs + (long)row * p.ctx.ts + (long)h * p.ctx.hs;
                const bf16* grow = p.dO.p + (long)b * p.dO.bs + (long)row * p.dO.ts + (long)h * p.dO.hs;
                for (int c0 = 0; c0 < dk; c0 += 8) {
                    float x[8], y[8];
                    Vec8<bf16>::load(orow + c0, x);
                    Vec8<bf16>::load(grow + c0, y);
#pragma unroll
                    for (int e = 0; e < 8; ++e) D = fmaf(x[e], y[e], D);
                }
                lse = p.lse[bh * p.T1p + row];
                p.Dvec[bh * p.T1p + row] = D;
            }
#pragma unroll 1
            for (int it = 0; it < nblk; ++it) {
                mbar_wait(x_full, (uint32_t)(it & 1));
                tcgen05_fence_after();
                const int rem = lim_row - it * BN;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t s[32], g[32];
                    tmem_ld32(tlane + half * 32, s);
                    tmem_ld32(tlane + 64 + half * 32, g);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float pr = ex2f(fmaf(__uint_as_float(s[c]), p.scale_log2, -lse));
                        const float ds = pr * (__uint_as_float(g[c]) - D) * p.scale;
                        v[c] = (c + half * 32 < rem) ? ds : 0.f;
                    }
                    store_w32(sW1, t, half * 32, v);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                signal_w();
            }
            bf16* qrow = p.out1.p + (long)b * p.out1.bs + (long)row * p.out1.ts + (long)h * p.out1.hs;
            if (nblk > 0) {
                mbar_wait(acc_full, 0);
                tcgen05_fence_after();
            }
#pragma unroll 1
            for (int c0 = 0; c0 < dk; c0 += 16) {
                uint32_t a[16];
                if (nblk > 0) {
                    tmem_ld16(tlane + 128 + c0, a);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) a[c] = 0u;
                }
                if (row_in) {
                    uint4 w0, w1;
                    w0.x = pack2(__uint_as_float(a[0]), __uint_as_float(a[1])); w0.y = pack2(__uint_as_float(a[2]), __uint_as_float(a[3]));
                    w0.z = pack2(__uint_as_float(a[4]), __uint_as_float(a[5])); w0.w = pack2(__uint_as_float(a[6]), __uint_as_float(a[7]));
                    w1.x = pack2(__uint_as_float(a[8]), __uint_as_float(a[9])); w1.y = pack2(__uint_as_float(a[10]), __uint_as_float(a[11]));
                    w1.z = pack2(__uint_as_float(a[12]), __uint_as_float(a[13])); w1.w = pack2(__uint_as_float(a[14]), __uint_as_float(a[15]));
                    *reinterpret_cast<uint4*>(qrow + c0) = w0;
                    *reinterpret_cast<uint4*>(qrow + c0 + 8) = w1;
                }
            }
        } else {
            // ---- MODE_BWD_KV: this thread owns key `row`; columns of X are the 64 queries of the streamed block
            const bool key_in = row < p.T2;
            const bool key_vis = row < klen;
            const float* lse_bh = p.lse + bh * p.T1p;
            const float* D_bh = p.Dvec + bh * p.T1p;
#pragma unroll 1
            for (int it = 0; it < nblk; ++it) {
                const int q0 = (blk0 + it) * BN;
                mbar_wait(x_full, (uint32_t)(it & 1));
                tcgen05_fence_after();
                // query q = q0 + c is attended by this key iff q < T1, key < klen and (not causal or key <= q)
                const int c_lo = p.causal ? max(row - q0, 0) : 0;          // first visible column
                const int c_hi = key_vis ? min(BN, p.T1 - q0) : 0;         // one past the last
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t s[32], g[32];
                    tmem_ld32(tlane + half * 32, s);
                    tmem_ld32(tlane + 64 + half * 32, g);
                    tmem_ld_wait();
                    float pv[32], dv[32];
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 l4 = __ldg(reinterpret_cast<const float4*>(lse_bh + q0 + half * 32) + c4);
                        const float4 d4 = __ldg(reinterpret_cast<const float4*>(D_bh + q0 + half * 32) + c4);
                        const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = c4 * 4 + e, cc = c + half * 32;
                            const bool vis = cc >= c_lo && cc < c_hi;
                            const float pr = ex2f(fmaf(__uint_as_float(s[c]), p.scale_log2, -ls[e]));
                            const float ds = pr * (__uint_as_float(g[c]) - dd[e]) * p.scale;
                            pv[c] = vis ? pr : 0.f;
                            dv[c] = vis ? ds : 0.f;
                        }
                    }
                    store_w32(sW1, t, half * 32, pv);
                    store_w32(sW2, t, half * 32, dv);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                signal_w();
            }
            bf16* vrow = p.out1.p + (long)b * p.out1.bs + (long)row * p.out1.ts + (long)h * p.out1.hs;
            bf16* krow = p.out2.p + (long)b * p.out2.bs + (long)row * p.out2.ts + (long)h * p.out2.hs;
            if (nblk > 0) {
                mbar_wait(acc_full, 0);
                tcgen05_fence_after();
            }
#pragma unroll 1
            for (int which = 0; which < 2; ++which) {
                bf16* orow = which ? krow : vrow;
                const uint32_t col = which ? colA2 : colA1;
#pragma unroll 1
                for (int c0 = 0; c0 < dk; c0 += 16) {
                    uint32_t a[16];
                    if (nblk > 0) {
                        tmem_ld16(tlane + col + c0, a);
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) a[c] = 0u;
                    }
                    if (key_in) {
                        uint4 w0, w1;
                        w0.x = pack2(__uint_as_float(a[0]), __uint_as_float(a[1])); w0.y = pack2(__uint_as_float(a[2]), __uint_as_float(a[3]));
                        w0.z = pack2(__uint_as_float(a[4]), __uint_as_float(a[5])); w0.w = pack2(__uint_as_float(a[6]), __uint_as_float(a[7]));
                        w1.x = pack2(__uint_as_float(a[8]), __uint_as_float(a[9])); w1.y = pack2(__uint_as_float(a[10]), __uint_as_float(a[11]));
                        w1.z = pack2(__uint_as_float(a[12]), __uint_as_float(a[13])); w1.w = pack2(__uint_as_float(a[14]), __uint_as_float(a[15]));
                        *reinterpret_cast<uint4*>(orow + c0) = w0;
                        *reinterpret_cast<uint4*>(orow + c0 + 8) = w1;
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

static size_t smem_bytes(int mode, int KA) {
    const bool bwd = mode != MODE_FWD;
    size_t n = (size_t)KA * ROW_PANEL * (bwd ? 2 : 1) + (size_t)NST * 2 * KA * BLK_PANEL + (size_t)BM * 128 * (mode == MODE_BWD_KV ? 2 : 1);
    return n + 1024 /*alignment slack*/ + 128 /*barriers*/;
}

static bool view_ok(const void* ptr, int64_t bs, int64_t ts, int64_t hs) {
    return ptr && (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && bs % 8 == 0 && ts % 8 == 0 && hs % 8 == 0;
}

// 4-D map {d_k, T, H, B} of a (B,T,H,d_k) strided bf16 view, box 64 x rows
static bool map_of(CUtensorMap* m, const void* ptr, int dk, int T, int H, int B, int64_t bs, int64_t ts, int64_t hs, int rows) {
    long dim[4] = {dk, T, H, B}, str[4] = {1, ts, hs, bs};
    return make_map(m, ptr, dim, str, 64, rows);
}

template <int MODE>
static int launch(const Params& p, int row_tiles, cudaStream_t st) {
    const size_t smem = smem_bytes(MODE, p.KA);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(attn_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    if (attr_err != cudaSuccess) return set_error(S2S_ERR_CUDA, "attn_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(attr_err));
    dim3 grid((unsigned)row_tiles, (unsigned)p.H, (unsigned)p.B);
    attn_tc_kernel<MODE><<<grid, NUM_THREADS, smem, st>>>(p);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

}  // namespace atc
}  // namespace s2s

using namespace s2s;

extern "C" int s2s_attn_fwd_tc(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, const void* v, int64_t kv_bs,
                               int64_t kv_ts, int64_t kv_hs, void* ctx, int64_t c_bs, int64_t c_ts, int64_t c_hs, float* lse, void* P,
                               int64_t ld, const int32_t* klens, int B, int H, int T1, int T2, int dk, float scale, int causal, void* stream) {
    using namespace atc;
    S2S_REQUIRE(q && k && v && ctx && lse && B > 0 && H > 0 && T1 > 0 && T2 > 0, "attn_fwd_tc: bad arguments");
    S2S_REQUIRE(dk >= 16 && dk <= 128 && dk % 16 == 0, "attn_fwd_tc: d_k %d must be a multiple of 16 in [16, 128]", dk);
    S2S_REQUIRE(B <= 65535 && H <= 65535, "attn_fwd_tc: too many batches / heads");
    S2S_REQUIRE(view_ok(q, q_bs, q_ts, q_hs) && view_ok(k, kv_bs, kv_ts, kv_hs) && view_ok(v, kv_bs, kv_ts, kv_hs) && view_ok(ctx, c_bs, c_ts, c_hs),
                "attn_fwd_tc: views must be 16-byte aligned with strides that are multiples of 8 elements");
    S2S_REQUIRE(!P || (ld >= T2 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0), "attn_fwd_tc: P rows must be 16-byte aligned (ld %% 8 == 0)");
    Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.T1 = T1; p.T2 = T2; p.dk = dk; p.KA = dk <= 64 ? 1 : 2; p.T1p = (T1 + 63) / 64 * 64;
    p.causal = causal; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f; p.klens = klens;
    p.ctx = View{(bf16*)ctx, c_bs, c_ts, c_hs};
    p.lse = lse; p.P = (bf16*)P; p.ld = ld;
    bool ok = map_of(&p.tmR1, q, dk, T1, H, B, q_bs, q_ts, q_hs, BM) && map_of(&p.tmC1, k, dk, T2, H, B, kv_bs, kv_ts, kv_hs, BN) &&
              map_of(&p.tmC2, v, dk, T2, H, B, kv_bs, kv_ts, kv_hs, BN);
    if (!ok) return set_error(S2S_ERR_UNSUPPORTED, "attn_fwd_tc: operands cannot be described to TMA");
    p.tmR2 = p.tmR1;
    return launch<MODE_FWD>(p, (T1 + BM - 1) / BM, (cudaStream_t)stream);
}

extern "C" int s2s_attn_bwd_tc(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, const void* v, int64_t kv_bs,
                               int64_t kv_ts, int64_t kv_hs, const void* ctx, const void* dctx, int64_t c_bs, int64_t c_ts, int64_t c_hs,
                               const float* lse, float* dvec, void* dq, int64_t dq_bs, int64_t dq_ts, int64_t dq_hs, void* dk_out, void* dv_out,
                               int64_t dkv_bs, int64_t dkv_ts, int64_t dkv_hs, const int32_t* klens, int B, int H, int T1, int T2, int dk,
                               float scale, int causal, void* stream) {
    using namespace atc;
    S2S_REQUIRE(q && k && v && ctx && dctx && lse && dvec && dq && dk_out && dv_out && B > 0 && H > 0 && T1 > 0 && T2 > 0, "attn_bwd_tc: bad arguments");
    S2S_REQUIRE(dk >= 16 && dk <= 128 && dk % 16 == 0, "attn_bwd_tc: d_k %d must be a multiple of 16 in [16, 128]", dk);
    S2S_REQUIRE(B <= 65535 && H <= 65535, "attn_bwd_tc: too many batches / heads");
    S2S_REQUIRE(view_ok(q, q_bs, q_ts, q_hs) && view_ok(k, kv_bs, kv_ts, kv_hs) && view_ok(v, kv_bs, kv_ts, kv_hs) && view_ok(ctx, c_bs, c_ts, c_hs) &&
                    view_ok(dctx, c_bs, c_ts, c_hs) && view_ok(dq, dq_bs, dq_ts, dq_hs) && view_ok(dk_out, dkv_bs, dkv_ts, dkv_hs) &&
                    view_ok(dv_out, dkv_bs, dkv_ts, dkv_hs),
                "attn_bwd_tc: views must be 16-byte aligned with strides that are multiples of 8 elements");
    Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.T1 = T1; p.T2 = T2; p.dk = dk; p.KA = dk <= 64 ? 1 : 2; p.T1p = (T1 + 63) / 64 * 64;
    p.causal = causal; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f; p.klens = klens;
    p.ctx = View{(bf16*)ctx, c_bs, c_ts, c_hs};
    p.dO = View{(bf16*)dctx, c_bs, c_ts, c_hs};
    p.lse = const_cast<float*>(lse); p.Dvec = dvec;
    cudaStream_t st = (cudaStream_t)stream;
    // ---- dQ (also writes D)
    p.out1 = View{(bf16*)dq, dq_bs, dq_ts, dq_hs};
    bool ok = map_of(&p.tmR1, q, dk, T1, H, B, q_bs, q_ts, q_hs, BM) && map_of(&p.tmR2, dctx, dk, T1, H, B, c_bs, c_ts, c_hs, BM) &&
              map_of(&p.tmC1, k, dk, T2, H, B, kv_bs, kv_ts, kv_hs, BN) && map_of(&p.tmC2, v, dk, T2, H, B, kv_bs, kv_ts, kv_hs, BN);
    if (!ok) return set_error(S2S_ERR_UNSUPPORTED, "attn_bwd_tc: operands cannot be described to TMA");
    int rc = launch<MODE_BWD_Q>(p, (T1 + BM - 1) / BM, st);
    if (rc != S2S_OK) return rc;
    // ---- dK, dV
    p.out1 = View{(bf16*)dv_out, dkv_bs, dkv_ts, dkv_hs};
    p.out2 = View{(bf16*)dk_out, dkv_bs, dkv_ts, dkv_hs};
    ok = map_of(&p.tmR1, k, dk, T2, H, B, kv_bs, kv_ts, kv_hs, BM) && map_of(&p.tmR2, v, dk, T2, H, B, kv_bs, kv_ts, kv_hs, BM) &&
         map_of(&p.tmC1, q, dk, T1, H, B, q_bs, q_ts, q_hs, BN) && map_of(&p.tmC2, dctx, dk, T1, H, B, c_bs, c_ts, c_hs, BN);
    if (!ok) return set_error(S2S_ERR_UNSUPPORTED, "attn_bwd_tc: operands cannot be described to TMA");
    return launch<MODE_BWD_KV>(p, (T2 + BM - 1) / BM, st);
}
