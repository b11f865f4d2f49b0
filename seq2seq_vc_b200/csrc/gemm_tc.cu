// bf16 tensor-core GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma (TMEM accumulators) ->
// tcgen05.ld epilogue.  Backs s2s_gemm(mode = 1).
//
// One persistent CTA per SM, 10 warps:
//   warp 0      TMA producer   (cp.async.bulk.tensor 4-D, 128B-swizzled boxes, mbarrier complete_tx)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16 with BN <= 256, commits to
//               mbarriers); also owns the TMEM allocation (2 accumulator stages x 256 fp32 columns)
//   warps 2-9   epilogue       (warp -> TMEM lane quarter x two 64-column chunks; each thread owns one
//               output row: tcgen05.ld 32x32b.x16 x4 in flight before one wait, bias broadcast by warp
//               shuffles, 16-byte row stores or vector red.add for fp32 accumulate-in-place outputs)
// Pipelines: STAGES-deep smem ring (full/empty mbarriers) between TMA and MMA, 2-deep TMEM ring
// (tmem_full/tmem_empty) between MMA and epilogue, static round-robin tile scheduler.
//
// Operands are described by 4-D TMA tensor maps built per call from the s2s_gemm_t strides, so the
// same kernel serves plain Linear layers, the head-strided attention views (B, H batches),
// transposed ("MN-major") operands of the weight-gradient GEMMs, and the `taps` convolution form
// (the tap index is a TMA coordinate: A rows shift by t, B selects weight slice t).
// K / M / N tails are handled by TMA zero fill; nothing is padded in HBM.
//
// Lessons baked into the structure (ncu + in-kernel clock traces, see profiles/ and DESIGN.md):
//  * CODE SIZE: the epilogue runs once per tile; when its straight-line body outgrew the
//    instruction cache (a fully unrolled, feature-complete version was 47 KB; the fp32 one 1.5 MB)
//    it ran at instruction-fetch speed (stall_no_inst) and was 2.5x slower than the MMA main loop.
//    The kernel is therefore specialised at compile time by epilogue kind (EPI_*), rare paths live
//    in one rolled non-inlined helper, and the tile decode is a non-inlined function.
//  * TMEM reads contend with the accumulator traffic of the next tile's MMAs: all loads of a
//    32 x 64 block are issued back to back before a single wait, and the accumulator stage is
//    released as soon as a warp's values are in registers.
//  * tcgen05.mma in cta_group::1 reads both operands from shared memory; 128 x 256 x 16 instructions
//    (BN = 256) halve the A re-reads per FLOP compared with 128 x 128.
//  * The main loop is paced by shared-memory traffic, not by the tensor pipe: one 128 x BN x 16 MMA took
//    70 + 0.68 BN cycles (245 @ BN = 256; the pipe's floor is 128), tracking the TMA writes + operand reads per
//    MMA.  A "tall" single-CTA tile (two M = 128 MMAs per B stage: fewer L2 bytes, same smem reads, no accumulator
//    double buffering) was measured and lost on every shape of the step except 8192^3 (+9 %).  CG = 2 is the fix
//    the hardware offers: a CTA pair (cluster of 2 on one TPC) issues ONE 256 x BN x 16 cta_group::2 MMA, each
//    CTA staging its own 128 rows of A and only HALF of the B tile, so smem writes and reads per FLOP drop by a
//    third while both CTAs keep their double-buffered 128 x BN accumulators.
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace s2s {

int gemm_simt(const s2s_gemm_t& g, cudaStream_t st);

namespace tc {

constexpr int BM = 128, BK = 64, MAX_BN = 256, STAGES = 4, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = MAX_BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TMEM_COLS = 512;      // two accumulator stages x 256 fp32 columns: the whole tensor memory of the SM
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int STG_WARP_BYTES = 32 * 128;   // epilogue staging: 32 rows x 64 bf16 columns per warp, 16-byte chunks XOR-swizzled by row
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_WARPS * STG_WARP_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "shared-memory budget of one CTA");

// epilogue specialisations (compile time, to keep every instantiation's code small)
constexpr int EPI_BF16_PLAIN = 0;   // bf16 C = alpha * acc + bias, optional relu
constexpr int EPI_BF16_FULL = 1;    // + dropout, residual, accumulate, row mask
constexpr int EPI_F32 = 2;          // fp32 C: red.add (accumulate in place / split-K) or plain store of alpha * acc (weight gradients)
constexpr int EPI_BF16_GATE = 3;    // bf16 C = (alpha * acc + bias) * r_scale where R > 0, else 0: ReLU' of the layer below in the dX GEMM
constexpr int EPI_F32_FULL = 4;     // fp32 C with the whole epilogue (bias, relu, dropout, residual / gate, accumulate, row mask):
                                    // the float32 activations of the fp32-accurate mode (gemm_split.cu)
// every kind is its own kernel: code that a launch never executes still costs instruction-cache footprint (folding the gate and
// the fp32 epilogue into EPI_BF16_FULL / EPI_F32 made every GEMM of the bf16 step 12 % slower, measured)

// n / d for n < 2^31 without the ~100-cycle integer-division sequence (the tile decode and the per-k-block tap split sit on
// the single-thread critical paths of the TMA and MMA warps): q = umulhi(n, mul) >> shr, magic numbers from the host
struct FastDiv {
    uint32_t d, mul, shr;
    __host__ void set(uint32_t div) {
        d = div; mul = 0; shr = 0;
        if (div > 1) {
            uint32_t lg = 0;
            while ((1ull << lg) < div) ++lg;                    // ceil(log2(d))
            const uint32_t pw = 31 + lg;
            mul = (uint32_t)(((1ull << pw) + div - 1) / div);
            shr = pw - 32;
        }
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (__umulhi(n, mul) >> shr); }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

struct Params {
    CUtensorMap tmA, tmB;
    int M, N, K, taps, batch1, batch2;
    int a_mn, b_mn;          // 1 = operand contiguous along M / N ("MN-major"), 0 = along K
    int BN;                  // N tile (multiple of 16, <= 256; multiple of 32 for CTA pairs)
    int cg;                  // 1 = one CTA per 128 x BN tile; 2 = CTA pair per 256 x BN tile (cta_group::2)
    int mt, nt, kb_per_tap, kb_total, splits;
    int kb_per_split;
    FastDiv d_splits, d_nt, d_mt, d_batch2, d_kbtap;
    void* C; int c_f32; long c_rs, c_bs1, c_bs2;
    const void* R;
    int r_gate;              // 0: x + R (residual); 1: R > 0 ? x * r_scale : 0 (ReLU' gate)
    float r_scale;
    const float* bias;
    float alpha;
    int relu, accumulate, atomic_out;
    int vec_ok;              // C (and R) rows are 16-byte aligned: vector loads / stores / red.add in the epilogue
    Dropout drop;
    int mask_period, mask_offset, mask_lo, mask_hi;
    long long* trace;        // optional (S2S_GEMM_TRACE builds): CTA 0 records clock64 stamps of pipeline milestones
    static constexpr bool kGrouped = false;
};

// Grouped launch: up to MAX_GROUPS independent weight-gradient GEMMs (fp32 C accumulated in place with red.add, both operands
// MN-major, no batch / taps) share ONE persistent launch -- the seven dW = dy^T x products of a decoder layer are 9-36 tiles
// each, so alone every one of them needs a ~24-way split-K (14 MB of red.add traffic for a 590 KB result) and pays its own
// launch + prologue; together they fill the machine with a 2-3-way split.  Items are numbered group by group.
constexpr int MAX_GROUPS = 8;
struct Group {
    CUtensorMap tmA, tmB;
    void* C;
    long c_rs;
    int M, N, K, nt, kb_total, splits, kb_per_split;
    uint32_t item0, items;      // first item of the group and their number (= tiles x splits)
    float alpha;
    FastDiv d_splits, d_nt;
};
struct ParamsG : Params {
    int ngroups;
    Group grp[MAX_GROUPS];
    static constexpr bool kGrouped = true;
};
static_assert(sizeof(ParamsG) <= 4096, "kernel parameter space");

// the problem an item belongs to (the launch's only problem, or its group)
struct Prob {
    const CUtensorMap *tmA, *tmB;
    void* C;
    long c_rs;
    int M, N, K, kb_per_tap;
    float alpha;
};

#ifdef S2S_GEMM_TRACE
__device__ __forceinline__ void stamp(const Params& p, int slot) {
    if (p.trace && blockIdx.x == 0) p.trace[slot] = clock64();
}
// cycles a role spent blocked on a barrier (CTA 0 only): slots 8 = MMA warp on full, 9 = MMA warp on tmem_empty,
// 10 = TMA producer on empty, 11 = k-blocks issued by CTA 0
#define TRACE_WAIT(slot, stmt)                                  \
    do {                                                        \
        const long long _t0 = clock64();                        \
        stmt;                                                   \
        if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) p.trace[slot] += clock64() - _t0; \
    } while (0)
#define TRACE_COUNT(slot) do { if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) p.trace[slot] += 1; } while (0)
#else
__device__ __forceinline__ void stamp(const Params&, int) {}
#define TRACE_WAIT(slot, stmt) stmt
#define TRACE_COUNT(slot) do {} while (0)
#endif

// ---------------------------------------------------------------------------------------------
// tile scheduler
// ---------------------------------------------------------------------------------------------
struct Item {
    int b1, b2, m0, n0, kb0, kb1;
};
template <typename P>
__device__ __noinline__ void decode_item(const P& p, uint32_t item, Item& it, Prob& pr) {
    if constexpr (P::kGrouped) {
        int g = 0;
#pragma unroll 1
        while (g + 1 < p.ngroups && item >= p.grp[g + 1].item0) ++g;
        const Group& G = p.grp[g];
        uint32_t tile, split, mtile, ntile;
        G.d_splits.divmod(item - G.item0, tile, split);
        G.d_nt.divmod(tile, mtile, ntile);
        it.b1 = it.b2 = 0;
        it.m0 = (int)mtile * BM;
        it.n0 = (int)ntile * p.BN;
        it.kb0 = (int)split * G.kb_per_split;
        it.kb1 = min(G.kb_total, it.kb0 + G.kb_per_split);
        pr.tmA = &G.tmA; pr.tmB = &G.tmB; pr.C = G.C; pr.c_rs = G.c_rs; pr.M = G.M; pr.N = G.N; pr.K = G.K;
        pr.kb_per_tap = G.kb_total; pr.alpha = G.alpha;
        return;
    } else {
        uint32_t tile, split, ntile, mtile, bz, b1, b2;
        p.d_splits.divmod(item, tile, split);
        p.d_nt.divmod(tile, tile, ntile);
        p.d_mt.divmod(tile, bz, mtile);
        p.d_batch2.divmod(bz, b1, b2);
        it.b1 = (int)b1;
        it.b2 = (int)b2;
        it.m0 = (int)mtile * BM * p.cg;
        it.n0 = (int)ntile * p.BN;
        it.kb0 = (int)split * p.kb_per_split;
        it.kb1 = min(p.kb_total, it.kb0 + p.kb_per_split);
        pr.tmA = &p.tmA; pr.tmB = &p.tmB; pr.C = p.C; pr.c_rs = p.c_rs; pr.M = p.M; pr.N = p.N; pr.K = p.K;
        pr.kb_per_tap = p.kb_per_tap; pr.alpha = p.alpha;
    }
}

// rare epilogue paths (column tails, unaligned rows, fp32 C with residual): rolled and out of line
template <typename TC>
__device__ __noinline__ void epilogue_scalar(const Params& p, const float* v, TC* __restrict__ dst, const TC* __restrict__ rsrc,
                                             int nvalid, bool row_ok) {
#pragma unroll 1
    for (int j = 0; j < nvalid; ++j) {
        float x = v[j];
        if (p.atomic_out) {
            if (row_ok) atomicAdd(reinterpret_cast<float*>(dst) + j, x);
            continue;
        }
        if (rsrc) {
            const float r = to_f<TC>(rsrc[j]);
            x = p.r_gate ? (r > 0.f ? x * p.r_scale : 0.f) : x + r;
        }
        if (p.accumulate) x += to_f<TC>(dst[j]);
        dst[j] = from_f<TC>(row_ok ? x : 0.f);
    }
}

__device__ __forceinline__ void add_bf16x8(float* v, const uint4& r) {
    const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[2 * e] += __low2float(r2[e]); v[2 * e + 1] += __high2float(r2[e]); }
}
// residual add or ReLU' gate with 8 bf16 values of R
__device__ __forceinline__ void combine_bf16x8(float* v, const uint4& r, int gate, float gscale) {
    const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r);
    if (gate) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[2 * e] = __low2float(r2[e]) > 0.f ? v[2 * e] * gscale : 0.f;
            v[2 * e + 1] = __high2float(r2[e]) > 0.f ? v[2 * e + 1] * gscale : 0.f;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[2 * e] += __low2float(r2[e]); v[2 * e + 1] += __high2float(r2[e]); }
    }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* v) {
    uint4 w;
    uint32_t* wp = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        wp[e] = *reinterpret_cast<uint32_t*>(&h);
    }
    return w;
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t total_items(const Params& p) { return (uint32_t)((long)p.batch1 * p.batch2 * p.mt * p.nt * p.splits); }
__device__ __forceinline__ uint32_t total_items(const ParamsG& p) { return p.grp[p.ngroups - 1].item0 + p.grp[p.ngroups - 1].items; }

template <int EPI, int CG, typename P = Params>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ P p) {
    constexpr bool F32OUT = (EPI == EPI_F32 || EPI == EPI_F32_FULL);
    constexpr bool FULLISH = (EPI == EPI_BF16_FULL || EPI == EPI_F32_FULL);      // dropout / residual / accumulate / row mask
    using TC = typename std::conditional<F32OUT, float, bf16>::type;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                    // SWIZZLE_128B tiles need 1024 B alignment
    const uint32_t bars = base + STAGES * STAGE_BYTES;               // full[STAGES], empty[STAGES], tfull[2], tempty[2]
    __shared__ uint32_t tmem_base_slot;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {      // descriptor fetch overlaps barrier init / TMEM allocation instead of delaying the first TMA load
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
    }
    // CTA pair: rank 0 (the leader) owns the full / tmem_empty barriers and issues the MMAs for both CTAs
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) stamp(p, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), EPI_WARPS * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer's barriers are initialised before anyone signals them
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    // Programmatic dependent launch (host: S2S_GEMM_PDL): everything above touches only this CTA's shared / tensor memory and
    // the kernel parameters, so it may run while the previous kernel of the stream drains; nothing below (TMA reads of A / B,
    // bias / residual reads, C writes) may.  A no-op when the launch carries no programmatic dependency.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) stamp(p, 1);

    const uint32_t total = total_items(p);
    const int BN = p.BN;
    const int BNH = BN / CG;                                  // B columns staged by this CTA
    const uint32_t worker = blockIdx.x / CG, n_workers = gridDim.x / CG;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            const int b_boxes = p.b_mn ? (BNH + 63) / 64 : 1;
            // bytes landing on the (leader's) full barrier per stage: both CTAs' A and B boxes
            const uint32_t tx_bytes = (uint32_t)CG * ((uint32_t)A_BYTES + (uint32_t)(p.b_mn ? b_boxes * 8192 : BNH * 128));
            int stage = 0;
            uint32_t phase = 0;
            Item it;
            Prob pr;
#pragma unroll 1
            for (uint32_t item = worker; item < total; item += n_workers) {
                decode_item(p, item, it, pr);
                const int am0 = it.m0 + (int)rank * BM, bn0 = it.n0 + (int)rank * BNH;
                int t = P::kGrouped ? 0 : (int)p.d_kbtap.div((uint32_t)it.kb0);
                int kk = (it.kb0 - t * pr.kb_per_tap) * BK;
#pragma unroll 1
                for (int kb = it.kb0; kb < it.kb1; ++kb, kk += BK) {
                    if (kk >= pr.kb_per_tap * BK) { kk = 0; ++t; }
                    TRACE_WAIT(10, mbar_wait(empty_bar(stage), phase ^ 1u));
                    const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    if (kb == it.kb0 && item == worker) stamp(p, 2);
                    const uint32_t fb = CG == 2 ? mapa_shared(full_bar(stage), 0) : full_bar(stage);
                    if (rank == 0) mbar_expect_tx(full_bar(stage), tx_bytes);
                    auto load = [&](uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3) {
                        if (CG == 2) tma_load_4d_pair(dst, map, fb, c0, c1, c2, c3);
                        else tma_load_4d(dst, map, fb, c0, c1, c2, c3);
                    };
                    if (p.a_mn) {
                        load(sa, pr.tmA, am0, kk, it.b2, it.b1);
                        load(sa + 8192, pr.tmA, am0 + 64, kk, it.b2, it.b1);
                    } else {
                        load(sa, pr.tmA, kk, am0 + t, it.b2, it.b1);
                    }
                    if (p.b_mn) {
#pragma unroll 1
                        for (int j = 0; j < b_boxes; ++j) load(sb + 8192 * j, pr.tmB, bn0 + 64 * j, kk, it.b2, it.b1);
                    } else {
                        load(sb, pr.tmB, kk, bn0, p.taps > 1 ? t : it.b2, it.b1);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ================= MMA issuer (leader CTA only for pairs: M = 256 spans both CTAs' A halves and accumulators) =================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
        // K-major: 16 bf16 = 32 B inside the 128 B swizzle row; MN-major: 16 k-rows of 128 B, 64-wide groups 8 KB apart
        const uint32_t a_step = p.a_mn ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;
        const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
        int stage = 0;
        uint32_t phase = 0;
        uint32_t n_items = 0;
        Item it;
        Prob pr;
#pragma unroll 1
        for (uint32_t item = worker; item < total; item += n_workers, ++n_items) {
            decode_item(p, item, it, pr);
            const int as = (int)(n_items & 1);
            const uint32_t aphase = (n_items >> 1) & 1u;
            TRACE_WAIT(9, mbar_wait(tempty_bar(as), aphase ^ 1u));
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * MAX_BN);
            int kk = (it.kb0 - (P::kGrouped ? 0 : (int)p.d_kbtap.div((uint32_t)it.kb0)) * pr.kb_per_tap) * BK;
#pragma unroll 1
            for (int kb = it.kb0; kb < it.kb1; ++kb, kk += BK) {
                if (kk >= pr.kb_per_tap * BK) kk = 0;
                const int ksteps = (min(BK, pr.K - kk) + UMMA_K - 1) / UMMA_K;
                TRACE_WAIT(8, mbar_wait(full_bar(stage), phase));
                TRACE_COUNT(11);
                tcgen05_fence_after();
                if (lane == 0 && n_items < 4 && kb == it.kb0) stamp(p, 16 + 8 * (int)n_items + 0);
                if (lane == 0 && n_items < 4 && kb == it.kb1 - 1) stamp(p, 16 + 8 * (int)n_items + 1);
                if (elect_one()) {
                    const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
                    const uint64_t ad0 = smem_desc(sa, a_lbo, 1024), bd0 = smem_desc(sb, b_lbo, 1024);
                    // the start-address field (bits 0-13, units of 16 B) advances; no carry out of it inside a stage
                    auto issue = [&](int k) {
                        const uint64_t ad = ad0 + (uint64_t)((k * a_step) >> 4), bd = bd0 + (uint64_t)((k * b_step) >> 4);
                        const uint32_t acc_flag = (kb > it.kb0 || k > 0) ? 1u : 0u;
                        if (CG == 2) umma_bf16_pair(tmem_d, ad, bd, idesc, acc_flag); else umma_bf16(tmem_d, ad, bd, idesc, acc_flag);
                    };
                    if (ksteps == BK / UMMA_K) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) issue(k);
                    } else {
#pragma unroll 1
                        for (int k = 0; k < ksteps; ++k) issue(k);
                    }
                    if (CG == 2) {
                        umma_commit_pair(empty_bar(stage));
                        if (kb == it.kb1 - 1) umma_commit_pair(tfull_bar(as));
                    } else {
                        umma_commit(empty_bar(stage));
                        if (kb == it.kb1 - 1) umma_commit(tfull_bar(as));
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 2) {
        // ================= epilogue: 8 warps; warp -> TMEM lane quarter x 64-column chunks {ch, ch + 2} =================
        Dropout drop = p.drop;
        if (FULLISH) dropout_resolve(drop);
        const int quarter = warp & 3;             // TMEM lanes this warp may access: 32 * (warp_id % 4) ...
        const int ch = (warp - 2) >> 2;           // owns 64-column chunks ch and ch + 2 of the tile
        // bf16 outputs leave through a per-warp staging buffer: each thread owns an output ROW (TMEM lane), so direct
        // 16-byte stores touch 32 different lines per instruction (the L1 store path paced the whole epilogue: ~5.7k
        // cycles per 128 x 256 tile); staged, every store instruction writes 4 rows x 128 contiguous bytes
        const uint32_t stg = bars + 256u + (uint32_t)(warp - 2) * STG_WARP_BYTES;
        uint32_t n_items = 0;
        Item it;
        Prob pr;
#pragma unroll 1
        for (uint32_t item = worker; item < total; item += n_workers, ++n_items) {
            decode_item(p, item, it, pr);
            const int as = (int)(n_items & 1);
            const uint32_t aphase = (n_items >> 1) & 1u;
            const int m = it.m0 + (int)rank * BM + quarter * 32 + lane;       // the output row this thread owns
            const bool row_in = m < pr.M;
            TC* Crow = reinterpret_cast<TC*>(pr.C) + it.b1 * p.c_bs1 + it.b2 * p.c_bs2 + (long)m * pr.c_rs;
            const TC* Rrow = (EPI == EPI_BF16_PLAIN || EPI == EPI_F32 || !p.R) ? nullptr
                                 : reinterpret_cast<const TC*>(p.R) + it.b1 * p.c_bs1 + it.b2 * p.c_bs2 + (long)m * pr.c_rs;
            const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * MAX_BN);
            bool row_ok = true;
            if (FULLISH && p.mask_period > 0) {
                const int ph = (m + p.mask_offset) % p.mask_period;
                row_ok = (ph >= p.mask_lo) && (ph < p.mask_hi);
            }
            // bias values (lane j holds columns j and j + 32 of each 64-column chunk) and, for the full epilogue,
            // the residual row segments of the first chunk are fetched before the accumulator is ready
            float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
            if (EPI != EPI_F32 && p.bias) {
                const int n0a = it.n0 + ch * 64, n0b = it.n0 + (ch + 2) * 64;
                if (n0a + lane < pr.N) b00 = p.bias[n0a + lane];
                if (n0a + lane + 32 < pr.N) b01 = p.bias[n0a + lane + 32];
                if (n0b + lane < pr.N) b10 = p.bias[n0b + lane];
                if (n0b + lane + 32 < pr.N) b11 = p.bias[n0b + lane + 32];
            }
            uint4 rr[8];
            const bool res_vec = (EPI == EPI_BF16_FULL || EPI == EPI_BF16_GATE) && Rrow && p.vec_ok && row_in;
            auto load_res = [&](int u) {
                const int c0 = (ch + 2 * u) * 64;
#pragma unroll
                for (int g8 = 0; g8 < 8; ++g8) {
                    rr[g8] = make_uint4(0u, 0u, 0u, 0u);
                    if (res_vec && c0 + g8 * 8 < BN && it.n0 + c0 + g8 * 8 + 8 <= pr.N)
                        rr[g8] = *reinterpret_cast<const uint4*>(Rrow + it.n0 + c0 + g8 * 8);
                }
            };
            // residual rows are prefetched per thread-owned row; a ReLU' gate is applied later, on the coalesced store side
            const bool pre_res = (EPI == EPI_BF16_FULL);
            if (pre_res) load_res(0);
            mbar_wait(tfull_bar(as), aphase);
            tcgen05_fence_after();
            if (threadIdx.x == 64 && n_items < 4) stamp(p, 16 + 8 * (int)n_items + 2);
#pragma unroll 1
            for (int u = 0; u < 2; ++u) {
                const int c0 = (ch + 2 * u) * 64;
                const int ncols = min(64, BN - c0);
                uint32_t acc[4][16];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q * 16 < ncols) tmem_ld16(tbase + c0 + q * 16, acc[q]);
                tmem_ld_wait();
                if (u == 0 && threadIdx.x == 64 && n_items < 4) stamp(p, 16 + 8 * (int)n_items + 6);
                if (u == 1) {
                    // every value of this warp is in registers: hand the accumulator stage back to the MMA warp
                    tcgen05_fence_before();
                    __syncwarp();
                    if (threadIdx.x == 64 && n_items < 4) stamp(p, 16 + 8 * (int)n_items + 3);
                    if (lane == 0) {
                        if (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(as), 0)); else mbar_arrive(tempty_bar(as));
                    }
                }
                if (ncols <= 0) continue;          // warp-uniform
                // warp-uniform: the whole 64-column chunk is inside N and C rows are 16-byte aligned
                const bool stage_chunk = !F32OUT && p.vec_ok && !p.accumulate && ncols == 64 && it.n0 + c0 + 64 <= pr.N &&
                                         !(EPI == EPI_BF16_FULL && drop.thresh != 0u && (pr.N & 15));   // group-wise dropout mask: 16-aligned row starts
                if (!F32OUT && stage_chunk) {
                    // ---- hot path: straight-line code, no per-group bounds / alignment decisions ----
                    const uint32_t rowaddr = stg + (uint32_t)lane * 128u;
                    const bool has_bias = p.bias != nullptr, relu = p.relu != 0;
                    const bool has_res = (EPI == EPI_BF16_FULL) && Rrow != nullptr;
                    const bool has_gate = (EPI == EPI_BF16_GATE) && Rrow != nullptr;
                    const bool has_drop = (EPI == EPI_BF16_FULL) && drop.thresh != 0u;
                    const uint64_t didx0 = (uint64_t)(((long)(it.b1 * p.batch2 + it.b2) * pr.M + m) * (long)pr.N + it.n0 + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float v[16];
                        if (has_bias) {
                            const float bsel = u ? ((q & 2) ? b11 : b10) : ((q & 2) ? b01 : b00);
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj)
                                v[jj] = fmaf(__uint_as_float(acc[q][jj]), pr.alpha, __shfl_sync(0xffffffffu, bsel, (q & 1) * 16 + jj));
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) v[jj] = __uint_as_float(acc[q][jj]) * pr.alpha;
                        }
                        if (relu) {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) v[jj] = fmaxf(v[jj], 0.f);
                        }
                        if (EPI == EPI_BF16_FULL) {
                            if (has_drop) {
                                // one seed per 16-element group (common.cuh: dropout_factors); didx0 is a multiple of 16 here (N % 16 == 0:
                                // see stage_chunk), so the 16 columns of this q are exactly one group
                                float mk[16];
                                dropout_factors<16>(drop, didx0 + (uint64_t)(q * 16), mk);
#pragma unroll
                                for (int jj = 0; jj < 16; ++jj) v[jj] *= mk[jj];
                            }
                            if (has_res) { add_bf16x8(v, rr[2 * q]); add_bf16x8(v + 8, rr[2 * q + 1]); }
                            if (!row_ok) {
#pragma unroll
                                for (int jj = 0; jj < 16; ++jj) v[jj] = 0.f;
                            }
                        }
                        if (EPI == EPI_BF16_GATE) {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) v[jj] *= p.r_scale;
                        }
                        st_shared_v4(rowaddr + (uint32_t)(((2 * q) ^ (lane & 7)) << 4), pack_bf16x8(v));
                        st_shared_v4(rowaddr + (uint32_t)(((2 * q + 1) ^ (lane & 7)) << 4), pack_bf16x8(v + 8));
                    }
                    if (pre_res && u == 0) load_res(1);
                    if (u == 0 && threadIdx.x == 64 && n_items < 4) stamp(p, 16 + 8 * (int)n_items + 7);
                    __syncwarp();
                    const int r_sub = lane >> 3, c16 = lane & 7;
                    const int row0 = it.m0 + (int)rank * BM + quarter * 32;
                    const long cofs = it.b1 * p.c_bs1 + it.b2 * p.c_bs2 + it.n0 + c0 + c16 * 8;
                    TC* Cblk = reinterpret_cast<TC*>(pr.C) + cofs;
                    if (EPI == EPI_BF16_GATE && has_gate) {
                        // ReLU' gate: the gate operand is read with the store's own coalesced pattern (4 rows x 128 B per instruction)
                        const TC* Gblk = reinterpret_cast<const TC*>(p.R) + cofs;
                        uint4 gq[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = i * 4 + r_sub;
                            gq[i] = make_uint4(0u, 0u, 0u, 0u);
                            if (row0 + r < pr.M) gq[i] = *reinterpret_cast<const uint4*>(Gblk + (long)(row0 + r) * pr.c_rs);
                        }
                        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = i * 4 + r_sub;
                            uint4 w = ld_shared_v4(stg + (uint32_t)r * 128u + (uint32_t)((c16 ^ (r & 7)) << 4));
                            __nv_bfloat162* w2 = reinterpret_cast<__nv_bfloat162*>(&w);
                            const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gq[i]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) w2[e] = __hmul2(w2[e], __hgt2(g2[e], zero2));
                            if (row0 + r < pr.M) *reinterpret_cast<uint4*>(Cblk + (long)(row0 + r) * pr.c_rs) = w;
                        }
                        __syncwarp();
                        continue;
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = i * 4 + r_sub;
                        const uint4 w = ld_shared_v4(stg + (uint32_t)r * 128u + (uint32_t)((c16 ^ (r & 7)) << 4));
                        if (row0 + r < pr.M) *reinterpret_cast<uint4*>(Cblk + (long)(row0 + r) * pr.c_rs) = w;
                    }
                    __syncwarp();
                    continue;
                }
                // ---- generic path: column / row tails, unaligned C, accumulate-in-place, fp32 outputs ----
                if (EPI == EPI_BF16_GATE) load_res(u);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q * 16 >= ncols) continue; // warp-uniform
                    float v[16];
                    if (EPI == EPI_F32) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) v[jj] = __uint_as_float(acc[q][jj]) * pr.alpha;
                    } else {
                        if (p.bias) {                  // warp-uniform
                            const float bsel = u ? ((q & 2) ? b11 : b10) : ((q & 2) ? b01 : b00);
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj)
                                v[jj] = fmaf(__uint_as_float(acc[q][jj]), pr.alpha, __shfl_sync(0xffffffffu, bsel, (q & 1) * 16 + jj));
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) v[jj] = __uint_as_float(acc[q][jj]) * pr.alpha;
                        }
                        if (p.relu) {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) v[jj] = fmaxf(v[jj], 0.f);
                        }
                    }
                    const int n_base = it.n0 + c0 + q * 16;
                    if (FULLISH && drop.thresh != 0u) {
                        const uint64_t didx = (uint64_t)(((long)(it.b1 * p.batch2 + it.b2) * pr.M + m) * (long)pr.N + n_base);
                        if ((didx & 15ull) == 0ull) {                // one seed per aligned group of 16 (common.cuh: dropout_factors)
                            float mk[16];
                            dropout_factors<16>(drop, didx, mk);
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) v[jj] *= mk[jj];
                        } else {
#pragma unroll 2
                            for (int jj = 0; jj < 16; ++jj) v[jj] *= dropout_factor(drop, didx + jj);
                        }
                    }
                    if (!row_in || n_base >= pr.N) continue;
                    TC* dst = Crow + n_base;
                    const bool fast = p.vec_ok && n_base + 16 <= pr.N;
                    if (EPI == EPI_F32) {
                        float* d32 = reinterpret_cast<float*>(dst);
                        if (fast && p.atomic_out) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) red_add_v4(d32 + 4 * e, v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        } else if (fast && !p.accumulate) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) *reinterpret_cast<float4*>(d32 + 4 * e) = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                        } else {
                            float vt[16];      // a COPY for the out-of-line helper: passing v[] itself would force it into local memory
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) vt[jj] = v[jj];
                            epilogue_scalar<TC>(p, vt, dst, nullptr, min(16, pr.N - n_base), true);
                        }
                    } else if (EPI == EPI_F32_FULL) {
                        float* d32 = reinterpret_cast<float*>(dst);
                        if (fast) {
                            // float32 activations (the fp32-accurate mode): residual / gate / accumulate with 16-byte accesses
                            const float* r32 = Rrow ? reinterpret_cast<const float*>(Rrow) + n_base : nullptr;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float4 x = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
                                if (r32) {
                                    const float4 r = *reinterpret_cast<const float4*>(r32 + 4 * e);
                                    if (p.r_gate) {
                                        x.x = r.x > 0.f ? x.x * p.r_scale : 0.f; x.y = r.y > 0.f ? x.y * p.r_scale : 0.f;
                                        x.z = r.z > 0.f ? x.z * p.r_scale : 0.f; x.w = r.w > 0.f ? x.w * p.r_scale : 0.f;
                                    } else {
                                        x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w;
                                    }
                                }
                                if (p.accumulate) {
                                    const float4 c = *reinterpret_cast<const float4*>(d32 + 4 * e);
                                    x.x += c.x; x.y += c.y; x.z += c.z; x.w += c.w;
                                }
                                *reinterpret_cast<float4*>(d32 + 4 * e) = row_ok ? x : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        } else {
                            float vt[16];
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) vt[jj] = v[jj];
                            epilogue_scalar<TC>(p, vt, dst, Rrow ? Rrow + n_base : nullptr, min(16, pr.N - n_base), row_ok);
                        }
                    } else if (fast) {
                        if (EPI == EPI_BF16_GATE && Rrow) { combine_bf16x8(v, rr[2 * q], 1, p.r_scale); combine_bf16x8(v + 8, rr[2 * q + 1], 1, p.r_scale); }
                        if (EPI == EPI_BF16_FULL) {
                            if (Rrow) { add_bf16x8(v, rr[2 * q]); add_bf16x8(v + 8, rr[2 * q + 1]); }
                            if (p.accumulate) {          // never staged (stage_chunk excludes it)
                                add_bf16x8(v, *reinterpret_cast<const uint4*>(dst));
                                add_bf16x8(v + 8, *reinterpret_cast<const uint4*>(dst + 8));
                            }
                            if (!row_ok) {
#pragma unroll
                                for (int jj = 0; jj < 16; ++jj) v[jj] = 0.f;
                            }
                        }
                        *reinterpret_cast<uint4*>(dst) = pack_bf16x8(v);
                        *reinterpret_cast<uint4*>(dst + 8) = pack_bf16x8(v + 8);
                    } else {
                        {   // the helper takes an address: hand it a COPY, or v[] itself is forced into local memory and every
                            // tile of the fast path pays 16 local stores per 16 columns (seen as STL in the ncu source view)
                            float vt[16];
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) vt[jj] = v[jj];
                            epilogue_scalar<TC>(p, vt, dst, Rrow ? Rrow + n_base : nullptr, min(16, pr.N - n_base), row_ok);
                        }
                    }
                }
                if (pre_res && u == 0) load_res(1);
            }
            if (threadIdx.x == 64 && n_items < 4) stamp(p, 16 + 8 * (int)n_items + 4);
        }
    }
    tcgen05_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();   // pairs: neither CTA leaves while the peer can still signal its barriers
    if (threadIdx.x == 0) stamp(p, 7);
    if (warp == 1) {
        tcgen05_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
// Tile choice.  N tile from a small cycle model of one CTA's work (numbers from in-kernel clock traces):
//   tile = fixed (~1500) + epilogue (~16 / column) + k-blocks x 4 MMAs x (70 + 0.68 BN) cycles,  total = waves(tiles) x tile.
// Wide tiles amortise the A re-reads of cta_group::1 MMAs; a problem with few tiles is cut finer so that all SMs work.
// Split-K candidates keep wide tiles (K is split instead).
// CTA pairs (cg = 2, 256 x BN tiles, half of B staged per CTA): measured launch-interleaved against single-CTA tiles
// (tools/gemm_probe.py ab, profiles/r01_gemm_pair_vs_single.txt): +5..9 % when the main loop dominates (K >= 768 and
// N >= 768: 49152 x 4608 x 1536 runs at 1186 vs 1091 TFLOP/s), -5..12 % on short-K / narrow-N shapes whose time is
// the epilogue and the per-tile fixed cost (16384 x 384 x 384: 20.5 vs 22.5 us) -- hence the gate below.
struct TileChoice { int bn, cg; };
static int g_force_cg = [] { const char* e = getenv("S2S_GEMM_CG"); return e ? atoi(e) : 0; }();   // 0 = gate below
static int g_force_bn = 0;                                                                            // 0 = cost model
static TileChoice pick_tile(int M, int N, long batches, long kblocks, bool splitk_candidate) {
    const long sms = num_sms();
    int cg = 1;
    const bool pair_ok = M > BM && N >= 32;
    if (g_force_cg == 2) cg = pair_ok ? 2 : 1;
    else if (g_force_cg == 0 && pair_ok && kblocks >= 12 && N >= 768 && ceil_div_l(M, 2 * BM) * batches * ceil_div_l(N, 256) >= sms / 4) cg = 2;
    const long mtiles = ceil_div_l(M, (long)BM * cg) * batches;
    const long workers = sms / cg;
    if (g_force_bn >= 16 * cg && g_force_bn <= MAX_BN && g_force_bn % (16 * cg) == 0) return TileChoice{g_force_bn, cg};
    TileChoice best{256, cg};
    double best_cost = -1.0;
    for (int bn = 256; bn >= 16 * cg; bn -= 16 * cg) {
        const long nt = (N + bn - 1) / bn;
        const long tiles = nt * mtiles;
        const long waves = splitk_candidate ? tiles : (tiles + workers - 1) / workers;
        const double mma = 4.0 * (70.0 + 0.68 * bn / cg);
        const double cost = (double)waves * (1500.0 + 16.0 * bn + (double)kblocks * mma);
        if (best_cost < 0 || cost < best_cost - 1e-9) { best_cost = cost; best.bn = bn; }
    }
    return best;
}

// Programmatic dependent launch of single-CTA GEMMs (experimental, off by default): the launch may begin while the previous
// kernel of the stream is still draining; the kernel's prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps
// that tail and griddepcontrol.wait orders every global-memory access after it.
static int g_pdl = [] { const char* e = getenv("S2S_GEMM_PDL"); return e ? atoi(e) : 0; }();
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
    const bool splitk_candidate = g.c_dtype == S2S_F32 && g.accumulate && !g.bias && !g.R && !g.relu && g.drop.p <= 0.f && g.mask_period == 0;
    const TileChoice tile = pick_tile(g.M, g.N, (long)g.batch1 * g.batch2, (long)ceil_div_l(g.K, BK) * g.taps, splitk_candidate);
    p.BN = tile.bn;
    p.cg = tile.cg;
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
                ok = make_map(&p.tmB, g.B, dim, str, BK, p.BN / p.cg);
            } else {
                long dim[4] = {g.K, g.N, g.batch2, g.batch1}, str[4] = {1, g.b_rs, g.b_bs2, g.b_bs1};
                ok = make_map(&p.tmB, g.B, dim, str, BK, p.BN / p.cg);
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
    p.mt = (int)ceil_div_l(g.M, (long)BM * p.cg);
    p.nt = (int)ceil_div_l(g.N, p.BN);
    p.kb_per_tap = (int)ceil_div_l(g.K, BK);
    p.kb_total = p.kb_per_tap * g.taps;
    p.C = g.C; p.c_f32 = (g.c_dtype == S2S_F32); p.c_rs = g.c_rs; p.c_bs1 = g.c_bs1; p.c_bs2 = g.c_bs2;
    p.R = g.R; p.r_gate = g.R ? g.r_mode : 0; p.r_scale = g.r_scale;
    p.bias = g.bias; p.alpha = g.alpha; p.relu = g.relu; p.accumulate = g.accumulate;
    p.drop = make_dropout(&g.drop);
    p.trace = g_trace;
    p.mask_period = g.mask_period; p.mask_offset = g.mask_offset; p.mask_lo = g.mask_lo; p.mask_hi = g.mask_hi;
    // split-K for skinny weight-gradient GEMMs: fp32 accumulate-in-place output, no other epilogue work
    const long tiles = (long)g.batch1 * g.batch2 * p.mt * p.nt;
    p.splits = 1;
    const bool plain = p.c_f32 && g.accumulate && !g.bias && !g.R && !g.relu && p.drop.thresh == 0u && g.mask_period == 0;
    if (plain) p.atomic_out = 1;          // fp32 accumulate-in-place: red.global.add, C is never read
    if (plain && tiles * 2 <= num_sms() / p.cg && p.kb_total >= 8) {
        long s = num_sms() / p.cg / tiles;
        long max_s = p.kb_total / 4;
        if (s > max_s) s = max_s;
        if (s > 1) p.splits = (int)s;
    }
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const int epb = p.c_f32 ? 4 : 8;
    p.vec_ok = (al16(g.C) && (g.R == nullptr || al16(g.R)) && g.c_rs % epb == 0 && g.c_bs1 % epb == 0 && g.c_bs2 % epb == 0) ? 1 : 0;
    // every split must own at least one k-block
    if (p.splits > 1) {
        int per = (p.kb_total + p.splits - 1) / p.splits;
        p.splits = (p.kb_total + per - 1) / per;
    }
    p.kb_per_split = (p.kb_total + p.splits - 1) / p.splits;
    p.d_splits.set((uint32_t)p.splits); p.d_nt.set((uint32_t)p.nt); p.d_mt.set((uint32_t)p.mt);
    p.d_batch2.set((uint32_t)p.batch2); p.d_kbtap.set((uint32_t)p.kb_per_tap);
    static const void* const kernels[2][5] = {
        {(const void*)gemm_tc_kernel<EPI_BF16_PLAIN, 1>, (const void*)gemm_tc_kernel<EPI_BF16_FULL, 1>, (const void*)gemm_tc_kernel<EPI_F32, 1>,
         (const void*)gemm_tc_kernel<EPI_BF16_GATE, 1>, (const void*)gemm_tc_kernel<EPI_F32_FULL, 1>},
        {(const void*)gemm_tc_kernel<EPI_BF16_PLAIN, 2>, (const void*)gemm_tc_kernel<EPI_BF16_FULL, 2>, (const void*)gemm_tc_kernel<EPI_F32, 2>,
         (const void*)gemm_tc_kernel<EPI_BF16_GATE, 2>, (const void*)gemm_tc_kernel<EPI_F32_FULL, 2>}};
    std::call_once(g_attr_once, [] {
        for (int c = 0; c < 2; ++c)
            for (int e = 0; e < 5; ++e)
                if (g_attr_err == cudaSuccess)
                    g_attr_err = cudaFuncSetAttribute(kernels[c][e], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    });
    if (g_attr_err != cudaSuccess) return set_error(S2S_ERR_CUDA, "gemm_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(g_attr_err));
    long items = tiles * p.splits;
    if (items >= (1L << 31)) return set_error(S2S_ERR_UNSUPPORTED, "gemm_tc: too many tiles");
    const long workers = num_sms() / p.cg;
    const unsigned grid = (unsigned)(items < workers ? items : workers) * (unsigned)p.cg;
    const bool plain_bf16 = !g.R && !g.accumulate && p.drop.thresh == 0u && g.mask_period == 0;
    const bool gate_only = g.R && p.r_gate && !g.accumulate && p.drop.thresh == 0u && g.mask_period == 0 && !g.relu;
    if (!p.c_f32 && g.R && p.r_gate && !gate_only)
        return set_error(S2S_ERR_UNSUPPORTED, "gemm_tc: a ReLU gate cannot be combined with dropout / accumulate / row mask / relu on bf16 outputs");
    const bool plain_f32 = !g.bias && !g.R && !g.relu && p.drop.thresh == 0u && g.mask_period == 0;      // weight gradients
    const int epi = p.c_f32 ? (plain_f32 ? EPI_F32 : EPI_F32_FULL) : (plain_bf16 ? EPI_BF16_PLAIN : (gate_only ? EPI_BF16_GATE : EPI_BF16_FULL));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (p.cg == 2) {       // CTA pair = thread-block cluster of 2 (placed on one TPC)
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    } else if (g_pdl) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    void* args[1] = {(void*)&p};
    const cudaError_t lerr = cudaLaunchKernelExC(&cfg, kernels[p.cg - 1][epi], args);
    if (lerr != cudaSuccess) return set_error(S2S_ERR_CUDA, "gemm_tc: launch failed: %s", cudaGetErrorString(lerr));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

// Grouped weight-gradient launch (see ParamsG).  Returns S2S_ERR_UNSUPPORTED when the set does not fit the grouped form (the
// caller then launches the GEMMs one by one).
int gemm_tc_grouped(const s2s_gemm_t* gs, int n, cudaStream_t st) {
    using namespace tc;
    if (n < 1 || n > MAX_GROUPS) return S2S_ERR_UNSUPPORTED;
    for (int i = 0; i < n; ++i) {
        const s2s_gemm_t& g = gs[i];
        const bool ok = g.a_dtype == S2S_BF16 && g.b_dtype == S2S_BF16 && g.c_dtype == S2S_F32 && g.accumulate && !g.bias && !g.R && !g.relu &&
                        g.drop.p <= 0.f && g.mask_period == 0 && g.batch1 * g.batch2 == 1 && g.taps == 1 && g.a_cs != 1 && g.b_cs != 1 &&
                        g.a_rs == 1 && g.b_rs == 1 && g.K > 1 && g.N >= 8 && g.M > 0 &&
                        (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && g.c_rs % 4 == 0;
        if (!ok) return S2S_ERR_UNSUPPORTED;
    }
    ParamsG p;
    memset(&p, 0, sizeof(p));
    // common N tile: the width that wastes the least padded work over the set
    int best_bn = 256;
    double best_waste = -1.0;
    for (int bn : {256, 192, 128}) {
        double w = 0.0;
        for (int i = 0; i < n; ++i) w += (double)ceil_div_l(gs[i].M, BM) * BM * ceil_div_l(gs[i].N, bn) * bn * (double)gs[i].K;
        if (best_waste < 0 || w < best_waste * 0.999) { best_waste = w; best_bn = bn; }
    }
    p.BN = best_bn; p.cg = 1; p.a_mn = 1; p.b_mn = 1; p.taps = 1; p.batch1 = p.batch2 = 1;
    p.c_f32 = 1; p.accumulate = 1; p.atomic_out = 1; p.vec_ok = 1; p.alpha = 1.f; p.trace = nullptr;
    p.ngroups = n;
    double work = 0.0;
    for (int i = 0; i < n; ++i) work += (double)ceil_div_l(gs[i].M, BM) * ceil_div_l(gs[i].N, p.BN) * ceil_div_l(gs[i].K, BK);
    const double target = work / (2.0 * num_sms()) > 8.0 ? work / (2.0 * num_sms()) : 8.0;      // k-blocks per item
    uint32_t item0 = 0;
    for (int i = 0; i < n; ++i) {
        const s2s_gemm_t& g = gs[i];
        Group& G = p.grp[i];
        long dimA[4] = {g.M, g.K, 1, 1}, strA[4] = {1, g.a_cs, 0, 0};
        long dimB[4] = {g.N, g.K, 1, 1}, strB[4] = {1, g.b_cs, 0, 0};
        if (!make_map(&G.tmA, g.A, dimA, strA, 64, BK) || !make_map(&G.tmB, g.B, dimB, strB, 64, BK)) return S2S_ERR_UNSUPPORTED;
        G.C = g.C; G.c_rs = g.c_rs; G.M = g.M; G.N = g.N; G.K = g.K; G.alpha = g.alpha;
        G.nt = (int)ceil_div_l(g.N, p.BN);
        G.kb_total = (int)ceil_div_l(g.K, BK);
        long sp = (long)(G.kb_total / target + 0.5);
        const long max_s = G.kb_total / 4;
        if (sp > max_s) sp = max_s;
        if (sp < 1) sp = 1;
        int per = (int)ceil_div_l(G.kb_total, sp);
        G.splits = (int)ceil_div_l(G.kb_total, per);          // every split owns at least one k-block
        G.kb_per_split = per;
        G.item0 = item0;
        G.items = (uint32_t)(ceil_div_l(g.M, BM) * G.nt * G.splits);
        item0 += G.items;
        G.d_splits.set((uint32_t)G.splits);
        G.d_nt.set((uint32_t)G.nt);
    }
    const void* kern = (const void*)gemm_tc_kernel<EPI_F32, 1, ParamsG>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [kern] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); });
    if (attr_err != cudaSuccess) return set_error(S2S_ERR_CUDA, "gemm_tc_grouped: cannot raise dynamic shared memory: %s", cudaGetErrorString(attr_err));
    const long workers = num_sms();
    const unsigned grid = (unsigned)(item0 < workers ? item0 : workers);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    void* args[1] = {(void*)&p};
    const cudaError_t lerr = cudaLaunchKernelExC(&cfg, kern, args);
    if (lerr != cudaSuccess) return set_error(S2S_ERR_CUDA, "gemm_tc_grouped: launch failed: %s", cudaGetErrorString(lerr));
    S2S_LAUNCH_OK();
    return S2S_OK;
}

long tc_fallback_count() { return g_tc_fallbacks; }

}  // namespace s2s

extern "C" void s2s_debug_gemm_trace(void* dev_buf8) { s2s::g_trace = (long long*)dev_buf8; }
extern "C" void s2s_debug_gemm_tile(int cg) {
    if (cg >= 16) { s2s::tc::g_force_bn = cg; return; }       // values >= 16 pin the N tile instead (cg choice untouched)
    s2s::tc::g_force_cg = (cg == 1 || cg == 2) ? cg : 0;
    if (cg == 0) s2s::tc::g_force_bn = 0;
}
extern "C" int64_t s2s_tc_fallback_count(void) { return (int64_t)s2s::tc_fallback_count(); }
