// Fused attention-probability kernels for small head dimensions (d_k <= 128, e.g. 48 / 64 / 96 of the VTN configs):
//
//   fwd :  P  = softmax_s( scale * Q K^T  masked by key length / causality )          (attention.py:95-104,76-85)
//   bwd :  dS = scale * P * (dP - sum_s P dP),  dP = dCtx V^T (+ dAtt)                (softmax' fused with its GEMM)
//
// With d_k = 48 the QK^T / dCtx V^T products are "all epilogue" for a 128 x 256 tcgen05 tile (K = 48 is three k-steps):
// the score matrix is written by the GEMM, re-read and re-written by the softmax, and the same again in backward.  Here
// the product is recomputed on the fly with warp-level mma.sync (bf16 in, fp32 accumulate) in two passes (row statistics,
// then output), so that the (B,H,T1,T2) matrix crosses HBM exactly once per direction (P write; P read + dS write).
// The kernel is HBM-write bound, not tensor bound, which is why the 5th-generation tensor path buys nothing here.
// One CTA = 4 warps = 64 query rows of one (b, h); K / V blocks of 64 keys are staged in shared memory.
#include "common.cuh"

namespace s2s {

struct AttnView {              // element (b, t, h, j) at base + b * bs + t * ts + h * hs + j
    const bf16* p;
    long bs, ts, hs;
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int AT_ROWS = 64;     // query rows per CTA (16 per warp)
constexpr int AT_KEYS = 64;     // keys per shared-memory block

// stage 64 keys x DK of K (or V) into shared memory, row pitch DK + 8 elements (conflict-free fragment reads)
template <int DK>
__device__ __forceinline__ void stage_keys(bf16* __restrict__ ks, const AttnView& kv, int b, int h, int key0, int T2) {
    constexpr int PITCH = DK + 8, CPR = DK / 8;
    for (int i = threadIdx.x; i < AT_KEYS * CPR; i += blockDim.x) {
        const int r = i / CPR, c = (i - r * CPR) * 8;
        const int key = key0 + r;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (key < T2) v = *reinterpret_cast<const uint4*>(kv.p + (long)b * kv.bs + (long)key * kv.ts + (long)h * kv.hs + c);
        *reinterpret_cast<uint4*>(ks + r * PITCH + c) = v;
    }
}

// same, with cp.async (16-byte, zero fill past T2): the copy of block n+1 overlaps the MMAs / exps of block n
template <int DK>
__device__ __forceinline__ void stage_keys_async(bf16* __restrict__ ks, const AttnView& kv, int b, int h, int key0, int T2) {
    constexpr int PITCH = DK + 8, CPR = DK / 8;
    for (int i = threadIdx.x; i < AT_KEYS * CPR; i += blockDim.x) {
        const int r = i / CPR, c = (i - r * CPR) * 8;
        const int key = key0 + r;
        const bf16* src = kv.p + (long)b * kv.bs + (long)(key < T2 ? key : 0) * kv.ts + (long)h * kv.hs + c;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ks + r * PITCH + c);
        const int nbytes = key < T2 ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(nbytes));
    }
    asm volatile("cp.async.commit_group;\n" ::);
}

// 16 x 64 block of (A K^T) for this warp: acc[n][4], n = 8-key tile
template <int DK>
__device__ __forceinline__ void block_scores(float (&acc)[8][4], const uint32_t (&afrag)[DK / 16][4], const bf16* __restrict__ ks) {
    constexpr int PITCH = DK + 8;
    const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        const bf16* kr = ks + (n * 8 + g) * PITCH + tig * 2;
#pragma unroll
        for (int kk = 0; kk < DK / 16; ++kk) {
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr + kk * 16);
            const uint32_t b1 = *reinterpret_cast<const uint32_t*>(kr + kk * 16 + 8);
            mma16816(acc[n], afrag[kk], b0, b1);
        }
    }
}

template <int DK>
__device__ __forceinline__ void load_afrag(uint32_t (&afrag)[DK / 16][4], const AttnView& q, int b, int h, int row_lo, int row_hi) {
    const int tig = threadIdx.x & 3;
    const bf16* r0 = q.p + (long)b * q.bs + (long)row_lo * q.ts + (long)h * q.hs + tig * 2;
    const bf16* r1 = q.p + (long)b * q.bs + (long)row_hi * q.ts + (long)h * q.hs + tig * 2;
#pragma unroll
    for (int kk = 0; kk < DK / 16; ++kk) {
        afrag[kk][0] = *reinterpret_cast<const uint32_t*>(r0 + kk * 16);
        afrag[kk][1] = *reinterpret_cast<const uint32_t*>(r1 + kk * 16);
        afrag[kk][2] = *reinterpret_cast<const uint32_t*>(r0 + kk * 16 + 8);
        afrag[kk][3] = *reinterpret_cast<const uint32_t*>(r1 + kk * 16 + 8);
    }
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// write this warp's 16 x 64 tile (fragment layout, already bf16-packed) through shared memory as 16-byte row chunks
__device__ __forceinline__ void store_tile(bf16* __restrict__ stg, const uint32_t (&lo)[8], const uint32_t (&hi)[8], bf16* __restrict__ out,
                                           long ld, int row0, int T1, int col0) {
    constexpr int SP = AT_KEYS + 8;            // staging pitch (elements): 144 B rows, conflict-free 4-byte writes
    const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    __syncwarp();
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        *reinterpret_cast<uint32_t*>(stg + g * SP + n * 8 + tig * 2) = lo[n];
        *reinterpret_cast<uint32_t*>(stg + (g + 8) * SP + n * 8 + tig * 2) = hi[n];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (lane >> 3) + i * 4, ch = lane & 7;
        const int row = row0 + r, col = col0 + ch * 8;
        if (row < T1 && col < ld) *reinterpret_cast<uint4*>(out + (long)row * ld + col) = *reinterpret_cast<const uint4*>(stg + r * SP + ch * 8);
    }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int DK>
__global__ void __launch_bounds__(128) attn_probs_fwd_kernel(AttnView q, AttnView k, bf16* __restrict__ P, const int32_t* __restrict__ klens,
                                                             int H, int T1, int T2, long ld, float scale, int causal) {
    constexpr int PITCH = DK + 8, SP = AT_KEYS + 8;
    __shared__ __align__(16) bf16 ks2[2][AT_KEYS * PITCH];
    __shared__ __align__(16) bf16 stg_all[4][16 * SP];
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const int row0 = blockIdx.x * AT_ROWS + warp * 16;
    const int r_lo = row0 + g, r_hi = row0 + g + 8;
    int klen = klens ? klens[b] : T2;
    klen = klen < 0 ? 0 : (klen > T2 ? T2 : klen);
    const int lim_lo = causal ? min(klen, r_lo + 1) : klen;
    const int lim_hi = causal ? min(klen, r_hi + 1) : klen;
    // keys any row of this CTA can see (later blocks are all-zero output)
    const int cta_last_row = min(T1, (int)(blockIdx.x + 1) * AT_ROWS) - 1;
    const int cta_lim = causal ? min(klen, cta_last_row + 1) : klen;
    uint32_t afrag[DK / 16][4];
    load_afrag<DK>(afrag, q, b, h, min(r_lo, T1 - 1), min(r_hi, T1 - 1));
    bf16* Pb = P + ((long)b * H + h) * (long)T1 * ld;
    bf16* stg = stg_all[warp];

    // Both passes walk the same nb key blocks; block seq+1 is copied (cp.async, double buffer) while block seq is consumed.
    //   pass 1 (seq < nb): online row max / sum          pass 2 (seq >= nb): probabilities, written once
    const int nb = (cta_lim + AT_KEYS - 1) / AT_KEYS;
    const int total = 2 * nb;
    float m_lo = -INFINITY, m_hi = -INFINITY, s_lo = 0.f, s_hi = 0.f, inv_lo = 0.f, inv_hi = 0.f;
    if (total > 0) stage_keys_async<DK>(ks2[0], k, b, h, 0, T2);
    for (int seq = 0; seq < total; ++seq) {
        const int blk = seq < nb ? seq : seq - nb;
        const int key0 = blk * AT_KEYS;
        if (seq + 1 < total) {
            const int nblk = (seq + 1 < nb) ? seq + 1 : seq + 1 - nb;
            stage_keys_async<DK>(ks2[(seq + 1) & 1], k, b, h, nblk * AT_KEYS, T2);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        float acc[8][4];
        block_scores<DK>(acc, afrag, ks2[seq & 1]);
        if (seq < nb) {
            float bm_lo = -INFINITY, bm_hi = -INFINITY;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int c = key0 + n * 8 + tig * 2;
                acc[n][0] = (c < lim_lo) ? acc[n][0] * scale : -INFINITY;
                acc[n][1] = (c + 1 < lim_lo) ? acc[n][1] * scale : -INFINITY;
                acc[n][2] = (c < lim_hi) ? acc[n][2] * scale : -INFINITY;
                acc[n][3] = (c + 1 < lim_hi) ? acc[n][3] * scale : -INFINITY;
                bm_lo = fmaxf(bm_lo, fmaxf(acc[n][0], acc[n][1]));
                bm_hi = fmaxf(bm_hi, fmaxf(acc[n][2], acc[n][3]));
            }
            bm_lo = quad_max(bm_lo);
            bm_hi = quad_max(bm_hi);
            const float nm_lo = fmaxf(m_lo, bm_lo), nm_hi = fmaxf(m_hi, bm_hi);
            float bs_lo = 0.f, bs_hi = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (nm_lo > -INFINITY) bs_lo += __expf(acc[n][0] - nm_lo) + __expf(acc[n][1] - nm_lo);
                if (nm_hi > -INFINITY) bs_hi += __expf(acc[n][2] - nm_hi) + __expf(acc[n][3] - nm_hi);
            }
            bs_lo = quad_sum(bs_lo);
            bs_hi = quad_sum(bs_hi);
            s_lo = (nm_lo > -INFINITY ? s_lo * __expf(m_lo - nm_lo) : 0.f) + bs_lo;
            s_hi = (nm_hi > -INFINITY ? s_hi * __expf(m_hi - nm_hi) : 0.f) + bs_hi;
            m_lo = nm_lo;
            m_hi = nm_hi;
            if (seq == nb - 1) { inv_lo = s_lo > 0.f ? 1.f / s_lo : 0.f; inv_hi = s_hi > 0.f ? 1.f / s_hi : 0.f; }
        } else {
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int c = key0 + n * 8 + tig * 2;
                const float p0 = (c < lim_lo) ? __expf(acc[n][0] * scale - m_lo) * inv_lo : 0.f;
                const float p1 = (c + 1 < lim_lo) ? __expf(acc[n][1] * scale - m_lo) * inv_lo : 0.f;
                const float p2 = (c < lim_hi) ? __expf(acc[n][2] * scale - m_hi) * inv_hi : 0.f;
                const float p3 = (c + 1 < lim_hi) ? __expf(acc[n][3] * scale - m_hi) * inv_hi : 0.f;
                lo[n] = pack_bf16(p0, p1);
                hi[n] = pack_bf16(p2, p3);
            }
            store_tile(stg, lo, hi, Pb, ld, row0, T1, key0);
        }
        __syncthreads();                                   // the buffer consumed here is refilled two iterations later
    }
    // key blocks no row of this CTA can see (causal / padded keys) and the columns up to ld: zeros, no compute
    {
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) { lo[n] = 0u; hi[n] = 0u; }
        for (int key0 = nb * AT_KEYS; key0 < (int)ld; key0 += AT_KEYS) store_tile(stg, lo, hi, Pb, ld, row0, T1, key0);
    }
}

// dS = scale * P * (dP - rowdot), dP = dCtx V^T (+ dAtt); P, dAtt, dS share the (B,H,T1,ld) layout.
// Pass 0 needs the whole row of P for the dot product, pass 1 needs it again: when `cache` is set the CTA keeps its
// 64 x ld slice of P in (dynamic) shared memory between the passes, so P is read from HBM once.
template <int DK>
__global__ void __launch_bounds__(128) attn_probs_bwd_kernel(AttnView dctx, AttnView v, const bf16* __restrict__ P, const bf16* __restrict__ dAtt,
                                                             bf16* __restrict__ dS, int H, int T1, int T2, long ld, float scale, int cache) {
    constexpr int PITCH = DK + 8, SP = AT_KEYS + 8;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ __align__(16) bf16 ks[AT_KEYS * PITCH];
    __shared__ __align__(16) bf16 stg_all[4][16 * SP];
    __shared__ __align__(16) bf16 stg2_all[4][16 * SP];
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const int row0 = blockIdx.x * AT_ROWS + warp * 16;
    uint32_t afrag[DK / 16][4];
    load_afrag<DK>(afrag, dctx, b, h, min(row0 + g, T1 - 1), min(row0 + g + 8, T1 - 1));
    const long base = ((long)b * H + h) * (long)T1 * ld;
    const bf16* Pb = P + base;
    const bf16* Ab = dAtt ? dAtt + base : nullptr;
    bf16* Sb = dS + base;
    bf16* stg = stg_all[warp];
    bf16* stg2 = stg2_all[warp];
    const int nblk = ((int)ld + AT_KEYS - 1) / AT_KEYS;
    const int cpitch = nblk * AT_KEYS + 8;                                  // cache row pitch (elements), 16-byte multiple
    bf16* pcache = reinterpret_cast<bf16*>(dyn_smem) + (long)warp * 16 * cpitch;

    // stage this warp's 16 x 64 tile of a (B,H,T1,ld) matrix into shared memory (coalesced 16-byte loads, zero fill)
    auto stage_rows = [&](bf16* dst, int pitch, const bf16* src, int col0) {
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = (lane >> 3) + i * 4, ch = lane & 7;
            const int row = row0 + r, col = col0 + ch * 8;
            uint4 val = make_uint4(0u, 0u, 0u, 0u);
            if (row < T1 && col < ld) val = *reinterpret_cast<const uint4*>(src + (long)row * ld + col);
            *reinterpret_cast<uint4*>(dst + r * pitch + ch * 8) = val;
        }
        __syncwarp();
    };
    auto frag = [&](const bf16* src, int pitch, int n, bool upper) -> float2 {
        const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(src + (g + (upper ? 8 : 0)) * pitch + n * 8 + tig * 2);
        return make_float2(__low2float(t), __high2float(t));
    };

    float dot_lo = 0.f, dot_hi = 0.f;
    for (int pass = 0; pass < 2; ++pass) {
        for (int key0 = 0; key0 < (int)ld; key0 += AT_KEYS) {
            __syncthreads();
            stage_keys<DK>(ks, v, b, h, key0, T2);
            __syncthreads();
            float acc[8][4];
            block_scores<DK>(acc, afrag, ks);
            const bf16* pt = stg;
            int ppitch = SP;
            if (cache) {
                pt = pcache + key0;
                ppitch = cpitch;
                if (pass == 0) stage_rows(pcache + key0, cpitch, Pb, key0);
            } else {
                stage_rows(stg, SP, Pb, key0);
            }
            if (Ab) stage_rows(stg2, SP, Ab, key0);
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int c = key0 + n * 8 + tig * 2;
                const float2 pl = frag(pt, ppitch, n, false), ph = frag(pt, ppitch, n, true);
                float d0 = acc[n][0], d1 = acc[n][1], d2 = acc[n][2], d3 = acc[n][3];
                if (Ab) {
                    const float2 al = frag(stg2, SP, n, false), ah = frag(stg2, SP, n, true);
                    d0 += al.x; d1 += al.y; d2 += ah.x; d3 += ah.y;
                }
                if (pass == 0) {
                    if (c < T2) { dot_lo += pl.x * d0; dot_hi += ph.x * d2; }
                    if (c + 1 < T2) { dot_lo += pl.y * d1; dot_hi += ph.y * d3; }
                } else {
                    const float o0 = (c < T2) ? scale * pl.x * (d0 - dot_lo) : 0.f;
                    const float o1 = (c + 1 < T2) ? scale * pl.y * (d1 - dot_lo) : 0.f;
                    const float o2 = (c < T2) ? scale * ph.x * (d2 - dot_hi) : 0.f;
                    const float o3 = (c + 1 < T2) ? scale * ph.y * (d3 - dot_hi) : 0.f;
                    lo[n] = pack_bf16(o0, o1);
                    hi[n] = pack_bf16(o2, o3);
                }
            }
            if (pass == 1) store_tile(stg, lo, hi, Sb, ld, row0, T1, key0);
        }
        if (pass == 0) { dot_lo = quad_sum(dot_lo); dot_hi = quad_sum(dot_hi); }
    }
}

static bool view_ok(const AttnView& a) {
    return (reinterpret_cast<uintptr_t>(a.p) & 15) == 0 && a.bs % 8 == 0 && a.ts % 8 == 0 && a.hs % 8 == 0;
}

}  // namespace s2s

using namespace s2s;

extern "C" int s2s_attn_probs_fwd(const void* q, int64_t q_bs, int64_t q_ts, int64_t q_hs, const void* k, int64_t k_bs, int64_t k_ts,
                                  int64_t k_hs, void* P, const int32_t* klens, int B, int H, int T1, int T2, int dk, int64_t ld,
                                  float scale, int causal, void* stream) {
    S2S_REQUIRE(q && k && P && B > 0 && H > 0 && T1 > 0 && T2 > 0 && ld >= T2, "attn_probs_fwd: bad arguments");
    S2S_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0, "attn_probs_fwd: P rows must be 16-byte aligned (ld %% 8 == 0)");
    S2S_REQUIRE(B <= 65535 && H <= 65535, "attn_probs_fwd: too many batches / heads");
    AttnView qv{(const bf16*)q, q_bs, q_ts, q_hs}, kv{(const bf16*)k, k_bs, k_ts, k_hs};
    S2S_REQUIRE(view_ok(qv) && view_ok(kv), "attn_probs_fwd: q / k strides must be multiples of 8 elements and 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div_l(T1, AT_ROWS), (unsigned)H, (unsigned)B);
#define S2S_AT_FWD(D) attn_probs_fwd_kernel<D><<<grid, 128, 0, st>>>(qv, kv, (bf16*)P, klens, H, T1, T2, ld, scale, causal)
    if (dk == 16) S2S_AT_FWD(16); else if (dk == 32) S2S_AT_FWD(32); else if (dk == 48) S2S_AT_FWD(48); else if (dk == 64) S2S_AT_FWD(64);
    else if (dk == 96) S2S_AT_FWD(96); else if (dk == 128) S2S_AT_FWD(128);
    else return set_error(S2S_ERR_UNSUPPORTED, "attn_probs_fwd: d_k %d not in {16,32,48,64,96,128}", dk);
#undef S2S_AT_FWD
    S2S_LAUNCH_OK();
    return S2S_OK;
}

extern "C" int s2s_attn_probs_bwd(const void* dctx, int64_t d_bs, int64_t d_ts, int64_t d_hs, const void* v, int64_t v_bs, int64_t v_ts,
                                  int64_t v_hs, const void* P, const void* dAtt, void* dS, int B, int H, int T1, int T2, int dk, int64_t ld,
                                  float scale, void* stream) {
    S2S_REQUIRE(dctx && v && P && dS && B > 0 && H > 0 && T1 > 0 && T2 > 0 && ld >= T2, "attn_probs_bwd: bad arguments");
    S2S_REQUIRE(ld % 8 == 0 && aligned16(P, dAtt, dS), "attn_probs_bwd: P / dAtt / dS rows must be 16-byte aligned (ld %% 8 == 0)");
    S2S_REQUIRE(B <= 65535 && H <= 65535, "attn_probs_bwd: too many batches / heads");
    AttnView dv{(const bf16*)dctx, d_bs, d_ts, d_hs}, vv{(const bf16*)v, v_bs, v_ts, v_hs};
    S2S_REQUIRE(view_ok(dv) && view_ok(vv), "attn_probs_bwd: dctx / v strides must be multiples of 8 elements and 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div_l(T1, AT_ROWS), (unsigned)H, (unsigned)B);
    // keep the CTA's 64 x ld slice of P in shared memory between the two passes when it leaves room for >= 2 CTAs per SM
    const long nblk = ceil_div_l(ld, AT_KEYS);
    const size_t cache_bytes = (size_t)4 * 16 * (nblk * AT_KEYS + 8) * sizeof(bf16);
    const int cache = cache_bytes <= 40 * 1024;        // beyond that the smaller number of resident CTAs costs more than the re-read
    const size_t dyn = cache ? cache_bytes : 0;
#define S2S_AT_BWD(D)                                                                                                          \
    do {                                                                                                                       \
        if (dyn > 0) /* static + dynamic shared memory can pass 48 KB; per-device attribute, set on every such call */          \
            S2S_CUDA_OK(cudaFuncSetAttribute(attn_probs_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024)); \
        attn_probs_bwd_kernel<D><<<grid, 128, dyn, st>>>(dv, vv, (const bf16*)P, (const bf16*)dAtt, (bf16*)dS, H, T1, T2, ld, scale, cache); \
    } while (0)
    if (dk == 16) S2S_AT_BWD(16); else if (dk == 32) S2S_AT_BWD(32); else if (dk == 48) S2S_AT_BWD(48); else if (dk == 64) S2S_AT_BWD(64);
    else if (dk == 96) S2S_AT_BWD(96); else if (dk == 128) S2S_AT_BWD(128);
    else return set_error(S2S_ERR_UNSUPPORTED, "attn_probs_bwd: d_k %d not in {16,32,48,64,96,128}", dk);
#undef S2S_AT_BWD
    S2S_LAUNCH_OK();
    return S2S_OK;
}
