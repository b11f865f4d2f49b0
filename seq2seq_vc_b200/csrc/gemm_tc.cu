// bf16 tensor-core GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma (TMEM accumulators) ->
// tcgen05.ld epilogue.  Backs s2s_gemm(mode = 1).
//
// One persistent CTA per SM, 6 warps:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled boxes, mbarrier complete_tx)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, commits to mbarriers);
//               also owns the TMEM allocation (2 accumulator stages x 128 fp32 columns)
//   warps 2-5   epilogue       (tcgen05.ld 32 lanes x 16 columns, fused alpha / bias / relu / dropout /
//               residual / accumulate / row-mask, vectorised global stores or fp32 red.add for split-K)
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) between TMA and MMA, 2-deep TMEM ring
// (tmem_full/tmem_empty) between MMA and epilogue, static round-robin tile scheduler.
//
// Operands are described by 4-D TMA tensor maps built per call from the s2s_gemm_t strides, so the
// same kernel serves plain Linear layers, the head-strided attention views (B, H batches),
// transposed ("MN-major") operands of the weight-gradient GEMMs, and the `taps` convolution form
// (the tap index is a TMA coordinate: A rows shift by t, B selects weight slice t).
// K / M / N tails are handled by TMA zero fill; nothing is padded in HBM.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace s2s {

int gemm_simt(const s2s_gemm_t& g, cudaStream_t st);

namespace tc {

constexpr int BM = 128, BK = 64, MAX_BN = 128, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = MAX_BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TMEM_COLS = 256;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
// epilogue staging: per epilogue warp 32 rows x 64 fp32 columns, row stride 68 floats (272 B) so that
// both the row-owner writes (16 B per lane, 32 rows) and the coalesced read-back are (nearly) conflict free
constexpr int STG_LD = 68, STG_BYTES = 32 * STG_LD * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES + EPI_WARPS * 64 * 4 + 1024 /*align slack*/ + 256 /*barriers*/;

struct Params {
    CUtensorMap tmA, tmB;
    int M, N, K, taps, batch1, batch2;
    int a_mn, b_mn;          // 1 = operand contiguous along M / N ("MN-major"), 0 = along K
    int BN;                  // N tile (multiple of 16, <= 128)
    int mt, nt, kb_per_tap, kb_total, splits;
    void* C; int c_f32; long c_rs, c_bs1, c_bs2;
    const void* R;
    const float* bias;
    float alpha;
    int relu, accumulate, atomic_out;
    int staged;              // bf16 C with 16 B-aligned rows: coalesced epilogue through shared memory
    Dropout drop;
    int mask_period, mask_offset, mask_lo, mask_hi;
    long long* trace;        // optional: CTA 0 records globaltimer stamps of pipeline milestones (debug)
};

__device__ __forceinline__ void stamp(const Params& p, int slot) {
    if (p.trace && blockIdx.x == 0) {
        p.trace[slot] = clock64();
    }
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (SWIZZLE_128B, sm_100 version field = 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
struct Item {
    int b1, b2, m0, n0, kb0, kb1;
};
__device__ __forceinline__ Item decode_item(const Params& p, long item_l) {
    Item it;
    const uint32_t item = (uint32_t)item_l;                 // host guarantees < 2^31 work items
    const uint32_t split = item % (uint32_t)p.splits;
    uint32_t tile = item / (uint32_t)p.splits;
    const uint32_t ntile = tile % (uint32_t)p.nt;
    tile /= (uint32_t)p.nt;
    const uint32_t mtile = tile % (uint32_t)p.mt;
    const uint32_t bz = tile / (uint32_t)p.mt;
    it.b1 = (int)(bz / (uint32_t)p.batch2);
    it.b2 = (int)(bz % (uint32_t)p.batch2);
    it.m0 = (int)mtile * BM;
    it.n0 = (int)ntile * p.BN;
    const int per = (p.kb_total + p.splits - 1) / p.splits;
    it.kb0 = (int)split * per;
    it.kb1 = min(p.kb_total, it.kb0 + per);
    return it;
}

// Code size matters here: the epilogue warps run this once per tile and a bloated body turns into
// instruction-cache misses (ncu: stall_no_inst dominated the first version).  Rare paths are kept in
// rolled loops / non-inlined helpers; only the 16-wide hot path is unrolled.

// direct (row-owner) store of 16 consecutive columns of one output row; used for fp32 C (weight
// gradients: red.add or plain stores) and for bf16 C whose rows are not 16 B aligned
template <typename TC>
__device__ __noinline__ void epilogue_direct(const Params& p, const Dropout& drop, float (&v)[16], TC* __restrict__ dst,
                                             const TC* __restrict__ rsrc, uint64_t didx, int nvalid, bool row_ok) {
    if (p.atomic_out) {
        if (!row_ok) return;
#pragma unroll 1
        for (int j = 0; j < nvalid; ++j) atomicAdd(reinterpret_cast<float*>(dst) + j, v[j]);
        return;
    }
#pragma unroll 1
    for (int j = 0; j < nvalid; ++j) {
        float x = v[j];
        if (p.relu) x = fmaxf(x, 0.f);
        x *= dropout_factor(drop, didx + j);
        if (rsrc) x += to_f<TC>(rsrc[j]);
        if (p.accumulate) x += to_f<TC>(dst[j]);
        dst[j] = from_f<TC>(row_ok ? x : 0.f);
    }
}

// N-tail of the coalesced path: fewer than 8 valid columns in this lane's group
template <typename TC>
__device__ __noinline__ void epilogue_tail(const Params& p, const float (&o)[8], TC* __restrict__ dst, const TC* __restrict__ rsrc,
                                           int nvalid, bool row_ok) {
#pragma unroll 1
    for (int q = 0; q < nvalid; ++q) {
        float x = o[q];
        if (rsrc) x += to_f<TC>(rsrc[q]);
        if (p.accumulate) x += to_f<TC>(dst[q]);
        dst[q] = from_f<TC>(row_ok ? x : 0.f);
    }
}

template <typename TC>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ Params p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                    // SWIZZLE_128B tiles need 1024 B alignment
    const uint32_t stg_base = base + STAGES * STAGE_BYTES;           // one staging tile per epilogue warp
    const uint32_t bias_base = stg_base + EPI_WARPS * STG_BYTES;     // 64 bias floats per epilogue warp
    const uint32_t bars = bias_base + EPI_WARPS * 64 * 4;            // full[STAGES], empty[STAGES], tfull[2], tempty[2]
    __shared__ uint32_t tmem_base_slot;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) stamp(p, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    if (threadIdx.x == 0) stamp(p, 1);

    const long total = (long)p.batch1 * p.batch2 * p.mt * p.nt * p.splits;
    const int BN = p.BN;
    const int b_boxes = p.b_mn ? (BN + 63) / 64 : 1;
    const uint32_t tx_bytes = (uint32_t)A_BYTES + (uint32_t)(p.b_mn ? b_boxes * 8192 : BN * 128);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long item = blockIdx.x; item < total; item += gridDim.x) {
                const Item it = decode_item(p, item);
                for (int kb = it.kb0; kb < it.kb1; ++kb) {
                    const int t = kb / p.kb_per_tap;
                    const int kk = (kb - t * p.kb_per_tap) * BK;
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    if (kb == it.kb0 && item == blockIdx.x) stamp(p, 2);
                    if (kb == it.kb0 && (item - blockIdx.x) / gridDim.x < 4) stamp(p, 16 + 8 * (int)((item - blockIdx.x) / gridDim.x) + 5);
                    mbar_expect_tx(full_bar(stage), tx_bytes);
                    if (p.a_mn) {
                        tma_load_4d(sa, &p.tmA, full_bar(stage), it.m0, kk, it.b2, it.b1);
                        tma_load_4d(sa + 8192, &p.tmA, full_bar(stage), it.m0 + 64, kk, it.b2, it.b1);
                    } else {
                        tma_load_4d(sa, &p.tmA, full_bar(stage), kk, it.m0 + t, it.b2, it.b1);
                    }
                    if (p.b_mn) {
                        for (int j = 0; j < b_boxes; ++j)
                            tma_load_4d(sb + 8192 * j, &p.tmB, full_bar(stage), it.n0 + 64 * j, kk, it.b2, it.b1);
                    } else {
                        tma_load_4d(sb, &p.tmB, full_bar(stage), kk, it.n0, p.taps > 1 ? t : it.b2, it.b1);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        int stage = 0;
        uint32_t phase = 0;
        long n_items = 0;
        for (long item = blockIdx.x; item < total; item += gridDim.x, ++n_items) {
            const Item it = decode_item(p, item);
            const int as = (int)(n_items & 1);
            const uint32_t aphase = (uint32_t)((n_items >> 1) & 1);
            mbar_wait(tempty_bar(as), aphase ^ 1u);
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * MAX_BN);
            for (int kb = it.kb0; kb < it.kb1; ++kb) {
                const int t = kb / p.kb_per_tap;
                const int kk = (kb - t * p.kb_per_tap) * BK;
                const int ksteps = (min(BK, p.K - kk) + UMMA_K - 1) / UMMA_K;
                mbar_wait(full_bar(stage), phase);
                tcgen05_fence_after();
                if (lane == 0 && n_items < 4 && kb == it.kb0) stamp(p, 16 + 8 * (int)n_items + 0);
                if (lane == 0 && n_items < 4 && kb == it.kb1 - 1) stamp(p, 16 + 8 * (int)n_items + 1);
                if (lane == 0) {
                    const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    for (int k = 0; k < ksteps; ++k) {
                        // K-major: 16 bf16 = 32 B inside the 128 B swizzle row; MN-major: 16 k-rows of 128 B
                        const uint64_t adesc = p.a_mn ? smem_desc(sa + k * 2048, 8192, 1024) : smem_desc(sa + k * 32, 16, 1024);
                        const uint64_t bdesc = p.b_mn ? smem_desc(sb + k * 2048, 8192, 1024) : smem_desc(sb + k * 32, 16, 1024);
                        umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));
                    if (kb == it.kb1 - 1) umma_commit(tfull_bar(as));
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ================= epilogue: 8 warps, warp -> (TMEM lane quarter, 64-column half) =================
        Dropout drop = p.drop;
        dropout_resolve(drop);
        const int ew = warp - 2;
        const int quarter = warp & 3;             // TMEM lanes this warp may access: 32 * (warp_id % 4) ...
        const int ch = ew >> 2;                   // which 64-column half of the tile
        unsigned char* smem_gen = smem_raw + (base - raw);
        float* stg = reinterpret_cast<float*>(smem_gen + STAGES * STAGE_BYTES + ew * STG_BYTES);
        float* bias_s = reinterpret_cast<float*>(smem_gen + STAGES * STAGE_BYTES + EPI_WARPS * STG_BYTES) + ew * 64;
        const int r4 = lane >> 3, cl = (lane & 7) * 8;   // coalesced mapping: 4 rows x 8 column groups of 8
        const bool staged = (sizeof(TC) == 2) && p.staged;
        const int ncols = min(64, BN - ch * 64);         // columns of this warp (<= 0: idle for this tile shape)
        long n_items = 0;
        for (long item = blockIdx.x; item < total; item += gridDim.x, ++n_items) {
            const Item it = decode_item(p, item);
            const int as = (int)(n_items & 1);
            const uint32_t aphase = (uint32_t)((n_items >> 1) & 1);
            const int row0 = it.m0 + quarter * 32;
            const int m = row0 + lane;                       // row owned while reading TMEM
            const int nw0 = it.n0 + ch * 64;                 // first column of this warp
            const long batch_lin = (long)(it.b1 * p.batch2 + it.b2) * p.M;
            TC* Cb = reinterpret_cast<TC*>(p.C) + it.b1 * p.c_bs1 + it.b2 * p.c_bs2;
            const TC* Rb = p.R ? reinterpret_cast<const TC*>(p.R) + it.b1 * p.c_bs1 + it.b2 * p.c_bs2 : nullptr;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * MAX_BN + ch * 64);
            // bias slice -> shared memory, residual rows -> registers (both hidden behind the MMA main loop)
            uint4 rr[8];
            if (ncols > 0) {
                bias_s[lane] = (p.bias && nw0 + lane < p.N) ? p.bias[nw0 + lane] : 0.f;
                bias_s[lane + 32] = (p.bias && nw0 + lane + 32 < p.N) ? p.bias[nw0 + lane + 32] : 0.f;
            }
            const bool pre = staged && Rb && cl < ncols && nw0 + cl + 8 <= p.N;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rr[i] = make_uint4(0u, 0u, 0u, 0u);
                if (pre && row0 + i * 4 + r4 < p.M)
                    rr[i] = *reinterpret_cast<const uint4*>(Rb + (long)(row0 + i * 4 + r4) * p.c_rs + nw0 + cl);
            }
            __syncwarp();
            mbar_wait(tfull_bar(as), aphase);
            tcgen05_fence_after();
            if (threadIdx.x == 64 && n_items < 4) stamp(p, 16 + 8 * (int)n_items + 2);
            // ---- all TMEM loads of this warp's 32 x 64 block are issued before the single wait
            uint32_t acc[4][16];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q * 16 < ncols) tmem_ld16(taddr + q * 16, acc[q]);
            tmem_ld_wait();
            // every value is in registers: hand the accumulator stage back to the MMA warp right away
            tcgen05_fence_before();
            __syncwarp();
            if (threadIdx.x == 64 && n_items < 4) stamp(p, 16 + 8 * (int)n_items + 3);
            if (lane == 0) mbar_arrive(tempty_bar(as));
            if (ncols <= 0) continue;
            // ---- alpha, bias, relu, dropout in the row-owner layout
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (q * 16 < ncols) {
                    float v[16];
                    const float4* b4 = reinterpret_cast<const float4*>(bias_s + q * 16);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float4 bb = b4[e];
                        v[4 * e + 0] = fmaf(__uint_as_float(acc[q][4 * e + 0]), p.alpha, bb.x);
                        v[4 * e + 1] = fmaf(__uint_as_float(acc[q][4 * e + 1]), p.alpha, bb.y);
                        v[4 * e + 2] = fmaf(__uint_as_float(acc[q][4 * e + 2]), p.alpha, bb.z);
                        v[4 * e + 3] = fmaf(__uint_as_float(acc[q][4 * e + 3]), p.alpha, bb.w);
                    }
                    const int n_base = nw0 + q * 16;
                    const uint64_t didx = (uint64_t)((batch_lin + m) * (long)p.N + n_base);
                    if (staged) {
                        if (p.relu) {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) v[jj] = fmaxf(v[jj], 0.f);
                        }
                        if (drop.thresh != 0u) {
#pragma unroll 4
                            for (int jj = 0; jj < 16; ++jj) v[jj] *= dropout_factor(drop, didx + jj);
                        }
                        float4* d4 = reinterpret_cast<float4*>(stg + lane * STG_LD + q * 16);
#pragma unroll
                        for (int e = 0; e < 4; ++e) d4[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                    } else if (m < p.M && n_base < p.N) {
                        bool row_ok = true;
                        if (p.mask_period > 0) {
                            const int ph = (m + p.mask_offset) % p.mask_period;
                            row_ok = (ph >= p.mask_lo) && (ph < p.mask_hi);
                        }
                        epilogue_direct<TC>(p, drop, v, Cb + (long)m * p.c_rs + n_base, Rb ? Rb + (long)m * p.c_rs + n_base : nullptr,
                                            didx, min(16, p.N - n_base), row_ok);
                    }
                }
            }
            if (!staged) continue;
            __syncwarp();
            // ---- coalesced read-back: each step covers 4 rows x (8 lanes x 8 columns), one 16 B store per lane
            const int n = nw0 + cl;
            const int nvalid = (cl < ncols) ? min(8, p.N - n) : 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rl = i * 4 + r4;
                const int row = row0 + rl;
                if (nvalid <= 0 || row >= p.M) continue;
                const float4 a0 = *reinterpret_cast<const float4*>(stg + rl * STG_LD + cl);
                const float4 a1 = *reinterpret_cast<const float4*>(stg + rl * STG_LD + cl + 4);
                float o[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                TC* dst = Cb + (long)row * p.c_rs + n;
                bool row_ok = true;
                if (p.mask_period > 0) {
                    const int ph = (row + p.mask_offset) % p.mask_period;
                    row_ok = (ph >= p.mask_lo) && (ph < p.mask_hi);
                }
                if (nvalid == 8) {
                    if (Rb) {
                        const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) { o[2 * q] += __low2float(r2[q]); o[2 * q + 1] += __high2float(r2[q]); }
                    }
                    if (p.accumulate) {
                        const uint4 cv = *reinterpret_cast<const uint4*>(dst);
                        const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&cv);
#pragma unroll
                        for (int q = 0; q < 4; ++q) { o[2 * q] += __low2float(c2[q]); o[2 * q + 1] += __high2float(c2[q]); }
                    }
                    uint4 w;
                    uint32_t* wp = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        __nv_bfloat162 hh = __floats2bfloat162_rn(row_ok ? o[2 * q] : 0.f, row_ok ? o[2 * q + 1] : 0.f);
                        wp[q] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(dst) = w;
                } else {
                    epilogue_tail<TC>(p, o, dst, Rb ? Rb + (long)row * p.c_rs + n : nullptr, nvalid, row_ok);
                }
            }
            if (threadIdx.x == 64 && n_items < 4) stamp(p, 16 + 8 * (int)n_items + 4);
            __syncwarp();     // the staging tile and bias slice are rewritten for the next tile
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp(p, 7);
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// dims/strides in elements, innermost first; stride[0] must be 1
static bool make_map(CUtensorMap* map, const void* base, const long (&dim)[4], const long (&stride)[4], int box0, int box1) {
    auto enc = get_encode();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
    cuuint64_t gdim[4], gstr[3];
    for (int i = 0; i < 4; ++i) {
        if (dim[i] <= 0 || dim[i] > 0xFFFFFFFFL) return false;
        gdim[i] = (cuuint64_t)dim[i];
    }
    long natural = dim[0];
    for (int i = 1; i < 4; ++i) {
        long s = stride[i];
        if (dim[i] == 1 || s <= 0) s = natural;       // size-1 (or broadcast over size 1) dims: any legal stride
        if (dim[i] > 1 && stride[i] <= 0) return false;
        if ((s * 2) % 16 != 0) {
            if (dim[i] == 1) s = (s + 7) / 8 * 8; else return false;
        }
        gstr[i - 1] = (cuuint64_t)s * 2;
        natural = s * dim[i];
    }
    cuuint32_t box[4] = {(cuuint32_t)box0, (cuuint32_t)box1, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static int pick_bn(int N) {
    int best = 128;
    long best_cost = -1;
    for (int bn = 128; bn >= 16; bn -= 16) {
        long tiles = (N + bn - 1) / bn;
        long cost = tiles * bn * 8 + tiles * 24;       // padded columns + a per-tile overhead term
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

static std::once_flag g_attr_once;
static cudaError_t g_attr_err = cudaSuccess;

}  // namespace tc

static long g_tc_fallbacks = 0;
static long long* g_trace = nullptr;

int gemm_tc(const s2s_gemm_t& g, cudaStream_t st) {
    using namespace tc;
    const bool dt_ok = g.a_dtype == S2S_BF16 && g.b_dtype == S2S_BF16;
    const bool taps_ok = g.taps == 1 || (g.a_cs == 1 && g.b_cs == 1 && g.batch1 * g.batch2 == 1);
    if (!dt_ok) return set_error(S2S_ERR_UNSUPPORTED, "gemm_tc: operands must be bf16 (a=%d b=%d)", g.a_dtype, g.b_dtype);
    Params p;
    memset(&p, 0, sizeof(p));
    bool ok = taps_ok && g.K > 0 && g.N >= 8;
    p.a_mn = (g.a_cs != 1) ? 1 : 0;      // contiguous along M
    p.b_mn = (g.b_cs != 1) ? 1 : 0;      // contiguous along N
    if (g.K == 1) { p.a_mn = (g.a_rs == 1); p.b_mn = (g.b_rs == 1); }
    p.BN = pick_bn(g.N);
    if (ok) {
        const long rowsA = (long)g.M + g.taps - 1;
        if (!p.a_mn) {
            long dim[4] = {g.K, rowsA, g.batch2, g.batch1}, str[4] = {1, g.a_rs, g.a_bs2, g.a_bs1};
            ok = make_map(&p.tmA, g.A, dim, str, BK, BM);
        } else {
            long dim[4] = {g.M, g.K, g.batch2, g.batch1}, str[4] = {1, g.a_cs, g.a_bs2, g.a_bs1};
            ok = make_map(&p.tmA, g.A, dim, str, 64, BK);
        }
    }
    if (ok) {
        if (!p.b_mn) {
            if (g.taps > 1) {
                long dim[4] = {g.K, g.N, g.taps, 1}, str[4] = {1, g.b_rs, g.b_ts, 0};
                ok = make_map(&p.tmB, g.B, dim, str, BK, p.BN);
            } else {
                long dim[4] = {g.K, g.N, g.batch2, g.batch1}, str[4] = {1, g.b_rs, g.b_bs2, g.b_bs1};
                ok = make_map(&p.tmB, g.B, dim, str, BK, p.BN);
            }
        } else {
            long dim[4] = {g.N, g.K, g.batch2, g.batch1}, str[4] = {1, g.b_cs, g.b_bs2, g.b_bs1};
            ok = make_map(&p.tmB, g.B, dim, str, 64, BK);
        }
    }
    if (!ok) {   // shapes TMA cannot describe (unaligned strides, N < 8, K == 0): CUDA-core kernel, counted
        ++g_tc_fallbacks;
        return gemm_simt(g, st);
    }
    p.M = g.M; p.N = g.N; p.K = g.K; p.taps = g.taps; p.batch1 = g.batch1; p.batch2 = g.batch2;
    p.mt = (int)ceil_div_l(g.M, BM);
    p.nt = (int)ceil_div_l(g.N, p.BN);
    p.kb_per_tap = (int)ceil_div_l(g.K, BK);
    p.kb_total = p.kb_per_tap * g.taps;
    p.C = g.C; p.c_f32 = (g.c_dtype == S2S_F32); p.c_rs = g.c_rs; p.c_bs1 = g.c_bs1; p.c_bs2 = g.c_bs2;
    p.R = g.R; p.bias = g.bias; p.alpha = g.alpha; p.relu = g.relu; p.accumulate = g.accumulate;
    p.drop = make_dropout(&g.drop);
    p.trace = g_trace;
    p.mask_period = g.mask_period; p.mask_offset = g.mask_offset; p.mask_lo = g.mask_lo; p.mask_hi = g.mask_hi;
    // split-K for skinny weight-gradient GEMMs: fp32 accumulate-in-place output, no other epilogue work
    const long tiles = (long)g.batch1 * g.batch2 * p.mt * p.nt;
    p.splits = 1;
    const bool plain = p.c_f32 && g.accumulate && !g.bias && !g.R && !g.relu && p.drop.thresh == 0u && g.mask_period == 0;
    if (plain) p.atomic_out = 1;          // fp32 accumulate-in-place: red.global.add, C is never read
    if (plain && tiles * 2 <= num_sms() && p.kb_total >= 8) {
        long s = num_sms() / tiles;
        long max_s = p.kb_total / 4;
        if (s > max_s) s = max_s;
        if (s > 1) p.splits = (int)s;
    }
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    p.staged = (!p.c_f32 && al16(g.C) && (g.R == nullptr || al16(g.R)) && g.c_rs % 8 == 0 && g.c_bs1 % 8 == 0 && g.c_bs2 % 8 == 0) ? 1 : 0;
    // every split must own at least one k-block
    if (p.splits > 1) {
        int per = (p.kb_total + p.splits - 1) / p.splits;
        p.splits = (p.kb_total + per - 1) / per;
    }
    std::call_once(g_attr_once, [] {
        g_attr_err = cudaFuncSetAttribute(gemm_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (g_attr_err == cudaSuccess)
            g_attr_err = cudaFuncSetAttribute(gemm_tc_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    });
    if (g_attr_err != cudaSuccess) return set_error(S2S_ERR_CUDA, "gemm_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(g_attr_err));
    long items = tiles * p.splits;
    if (items >= (1L << 31)) return set_error(S2S_ERR_UNSUPPORTED, "gemm_tc: too many tiles");
    unsigned grid = (unsigned)(items < num_sms() ? items : num_sms());
    if (p.c_f32) gemm_tc_kernel<float><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
    else gemm_tc_kernel<bf16><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

long tc_fallback_count() { return g_tc_fallbacks; }

}  // namespace s2s

extern "C" void s2s_debug_gemm_trace(void* dev_buf8) { s2s::g_trace = (long long*)dev_buf8; }
extern "C" int64_t s2s_tc_fallback_count(void) { return (int64_t)s2s::tc_fallback_count(); }
