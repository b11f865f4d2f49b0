// Monotonic alignment search on the GPU, bit-exact with the reference's numba routine
// (modules/alignments.py:63-93) and viterbi_decode's bincount / bin-loss (:281-310).
//
// One CTA per utterance; thread i owns text position i and carries Q[i, j-1] in a register while
// the CTA sweeps the mel axis j sequentially.  Each step needs only the left neighbour's previous
// value, exchanged by warp shuffle (plus one shared-memory slot per warp boundary), so a step costs
// one __syncthreads.  The backtrack decision "Q[i-1, j] >= Q[i, j]" is recorded as a ballot bitmask
// per (j, warp) in the workspace, so the backward walk never re-reads Q.  Arithmetic follows the
// reference exactly: row 0 is a sequential float32 prefix sum widened to float64, everything else is
// float64 max/add (no fused multiply-add is possible in this recurrence), ties go back.
#include "common.cuh"

namespace s2s {

constexpr int MAS_PF = 8;  // log_p prefetch depth (steps)

__global__ void mas_kernel(const float* __restrict__ log_p, const int32_t* __restrict__ text_lens,
                           const int32_t* __restrict__ feats_lens, int B, int T_feats, int T_text,
                           int32_t* __restrict__ paths, float* __restrict__ ds, float* bin_loss,
                           float* __restrict__ d_log_p, uint32_t* __restrict__ ws, int ws_words_per_utt) {
    extern __shared__ unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int i = threadIdx.x, lane = i & 31, warp = i >> 5, nwarps = blockDim.x >> 5;
    int t_mel = feats_lens[b], t_inp = text_lens[b];
    if (t_mel > T_feats) t_mel = T_feats;
    if (t_inp > T_text) t_inp = T_text;
    const float* lp = log_p + (size_t)b * T_feats * T_text;
    int32_t* path = paths + (size_t)b * T_feats;
    float* dsb = ds + (size_t)b * T_text;
    for (int j = i; j < T_feats; j += blockDim.x) path[j] = -1;
    for (int k = i; k < T_text; k += blockDim.x) dsb[k] = 0.f;
    if (t_mel <= 0 || t_inp <= 0) return;

    double* edge = reinterpret_cast<double*>(smem_raw);             // [2][nwarps] boundary values
    int* s_path = reinterpret_cast<int*>(edge + 2 * nwarps);        // [t_mel]
    __shared__ float red[32];
    uint32_t* bits = ws + (size_t)b * ws_words_per_utt;             // [T_feats][nwarps]

    const bool active = i < t_inp;
    const double NEG = -INFINITY;
    // column j = 0: only i == 0 is finite
    float acc32 = 0.f;
    double q = NEG;
    float cur[MAS_PF], nxt[MAS_PF];
#pragma unroll
    for (int u = 0; u < MAS_PF; ++u) cur[u] = (active && u < t_mel) ? lp[(size_t)u * T_text + i] : 0.f;
    if (i == 0) { acc32 = acc32 + cur[0]; q = (double)acc32; }

    for (int j0 = 0; j0 < t_mel; j0 += MAS_PF) {
#pragma unroll
        for (int u = 0; u < MAS_PF; ++u) {
            int jn = j0 + MAS_PF + u;
            nxt[u] = (active && jn < t_mel) ? lp[(size_t)jn * T_text + i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < MAS_PF; ++u) {
            const int j = j0 + u;  // q currently holds column j
            if (j < t_mel) {       // uniform across the CTA
                // neighbour value Q[i-1, j]
                double up = __shfl_up_sync(0xffffffffu, q, 1);
                const int par = j & 1;
                if (lane == 31) edge[par * nwarps + warp] = q;
                __syncthreads();
                if (lane == 0) up = (warp > 0) ? edge[par * nwarps + warp - 1] : NEG;
                // backtrack decision for column j: go to i-1 iff Q[i-1, j] >= Q[i, j]
                unsigned m = __ballot_sync(0xffffffffu, active && i > 0 && up >= q);
                if (lane == 0) bits[(size_t)j * nwarps + warp] = m;
                // advance to column j + 1
                if (j + 1 < t_mel) {
                    float l = (u + 1 < MAS_PF) ? cur[(u + 1) % MAS_PF] : nxt[0];
                    if (i == 0) {
                        acc32 = acc32 + l;
                        q = (double)acc32;
                    } else if (active && i < j + 2) {  // i < min(j' + 1, t_inp) with j' = j + 1
                        double mx = (up > q) ? up : q;
                        q = mx + (double)l;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < MAS_PF; ++u) cur[u] = nxt[u];
    }
    __threadfence_block();
    __syncthreads();
    if (i == 0) {
        int a = t_inp - 1;
        s_path[t_mel - 1] = a;
        for (int j = t_mel - 2; j >= 0; --j) {
            if (a != 0) {
                uint32_t m = bits[(size_t)j * nwarps + (a >> 5)];
                if ((m >> (a & 31)) & 1u) a -= 1;
            }
            s_path[j] = a;
        }
    }
    __syncthreads();
    float s = 0.f;
    const float gscale = -1.f / ((float)t_mel * (float)B);
    for (int j = i; j < t_mel; j += blockDim.x) {
        int a = s_path[j];
        path[j] = a;
        atomicAdd(&dsb[a], 1.f);
        s += lp[(size_t)j * T_text + a];
        if (d_log_p) d_log_p[((size_t)b * T_feats + j) * T_text + a] = gscale;
    }
    s = block_sum(s, red);
    if (i == 0 && bin_loss) atomicAdd(bin_loss, s * gscale);
}

}  // namespace s2s

using namespace s2s;

static int mas_block(int T_text) { int t = ((T_text + 31) / 32) * 32; return t < 32 ? 32 : t; }

extern "C" size_t s2s_mas_workspace_bytes(int B, int T_feats, int T_text) {
    if (B <= 0 || T_feats <= 0 || T_text <= 0) return 0;
    return (size_t)B * T_feats * (mas_block(T_text) / 32) * sizeof(uint32_t);
}

extern "C" int s2s_mas(const float* log_p, const int32_t* text_lens, const int32_t* feats_lens, int B, int T_feats,
                       int T_text, int32_t* paths, float* ds, float* bin_loss, float* d_log_p, void* workspace,
                       size_t workspace_bytes, void* stream) {
    S2S_REQUIRE(log_p && text_lens && feats_lens && paths && ds && workspace, "mas: null pointer");
    S2S_REQUIRE(B > 0 && T_feats > 0 && T_text > 0, "mas: bad shape");
    S2S_REQUIRE(T_text <= 1024, "mas: T_text %d > 1024 unsupported", T_text);
    S2S_REQUIRE(workspace_bytes >= s2s_mas_workspace_bytes(B, T_feats, T_text), "mas: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int block = mas_block(T_text), nw = block / 32;
    size_t smem = (size_t)2 * nw * sizeof(double) + (size_t)T_feats * sizeof(int);
    S2S_REQUIRE(smem <= 200 * 1024, "mas: T_feats %d too long for shared-memory path buffer", T_feats);
    if (smem > 48 * 1024)      // per-device attribute: set on every call that needs it (cheap, thread-safe)
        S2S_CUDA_OK(cudaFuncSetAttribute(mas_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (bin_loss) S2S_CUDA_OK(cudaMemsetAsync(bin_loss, 0, sizeof(float), st));
    mas_kernel<<<B, block, smem, st>>>(log_p, text_lens, feats_lens, B, T_feats, T_text, paths, ds, bin_loss, d_log_p,
                                       (uint32_t*)workspace, T_feats * nw);
    S2S_LAUNCH_OK();
    return S2S_OK;
}
