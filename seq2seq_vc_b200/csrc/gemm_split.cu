// fp32-accurate GEMM on the bf16 tensor cores: s2s_gemm(mode = 2).
//
// A float32 value x is split into bf16 pieces x0 = bf16(x), x1 = bf16(x - x0) (, x2 = bf16(x - x0 - x1)): two pieces carry
// 16 significand bits, three carry all 24.  The product a b is then the sum of the piece products that matter:
//     split_terms = 3:  a0 b0 + a1 b0 + a0 b1                                  (relative error ~ 2^-16 per product)
//     split_terms = 6:  a0 b0 + a0 b1 + a1 b0 + a1 b1 + a0 b2 + a2 b0          (~ 2^-23: float32-grade)
// each a bf16 x bf16 product accumulated exactly in the fp32 TMEM accumulator.  Instead of issuing several GEMMs the pieces
// are laid side by side along K in a workspace -- A' = [a0 | a1 | a0], B' = [b0 | b0 | b1] (and the six-term analogue) --
// so ONE launch of the tcgen05 kernel of gemm_tc.cu with K' = terms * K computes the sum, epilogue included.  The split
// kernels read the operands through their element strides (plain, transposed, head-strided, `taps` forms alike) and write
// dense K-major rows, K padded to a multiple of 8 for TMA.
#include "common.cuh"

namespace s2s {

int gemm_tc(const s2s_gemm_t& g, cudaStream_t st);

namespace split {

// piece pattern per operand: index of the piece that goes into K-segment s
__constant__ int kPatA3[3] = {0, 1, 0}, kPatB3[3] = {0, 0, 1};
__constant__ int kPatA6[6] = {0, 0, 1, 1, 0, 2}, kPatB6[6] = {0, 1, 0, 1, 2, 0};

struct Src {
    const float* p;
    long rs, cs, bs1, bs2;   // element strides of (row, k, batch1, batch2)
    int rows, K, Kp, terms, is_b;
    int nb2;                 // batch2 extent (batch index = b1 * nb2 + b2)
};

__device__ __forceinline__ void pieces(float x, bf16 (&pc)[3]) {
    pc[0] = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(pc[0]);
    pc[1] = __float2bfloat16_rn(r1);
    pc[2] = __float2bfloat16_rn(r1 - __bfloat162float(pc[1]));
}

// one thread per (row, k); threads run along k when the source is k-contiguous, along rows otherwise (coalesced reads
// either way; the strided side goes through a 32 x 33 shared-memory tile)
__global__ void __launch_bounds__(256) split_kernel(Src s, bf16* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int bz = blockIdx.z;
    const int b1 = bz / s.nb2, b2 = bz % s.nb2;
    const float* src = s.p + b1 * s.bs1 + b2 * s.bs2;
    bf16* out = dst + (long)bz * s.rows * ((long)s.terms * s.Kp);
    const int k0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
    const bool k_contig = (s.cs == 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = ty + i * 8;
        // k-contiguous: tile[row a][k tx]; row-contiguous: tile[k a][row tx] read, stored transposed
        const int r = k_contig ? r0 + a : r0 + tx;
        const int k = k_contig ? k0 + tx : k0 + a;
        float v = 0.f;
        if (r < s.rows && k < s.K) v = src[(long)r * s.rs + (long)k * s.cs];
        if (k_contig) tile[a][tx] = v; else tile[tx][a] = v;
    }
    __syncthreads();
    const int* pat = s.terms == 3 ? (s.is_b ? kPatB3 : kPatA3) : (s.is_b ? kPatB6 : kPatA6);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + ty + i * 8, k = k0 + tx;
        if (r < s.rows && k < s.Kp) {
            bf16 pc[3];
            pieces(tile[ty + i * 8][tx], pc);       // zero beyond K: the pad columns contribute nothing
            bf16* o = out + (long)r * ((long)s.terms * s.Kp) + k;
            for (int t = 0; t < s.terms; ++t) o[(long)t * s.Kp] = pc[pat[t]];
        }
    }
}

static int launch_split(const Src& s, bf16* dst, int nbatch, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div_l(s.Kp, 32), (unsigned)ceil_div_l(s.rows, 32), (unsigned)nbatch);
    if (grid.y > 65535 || grid.z > 65535) return set_error(S2S_ERR_UNSUPPORTED, "gemm(mode 2): operand too large for the split kernel grid");
    split_kernel<<<grid, 256, 0, st>>>(s, dst);
    S2S_LAUNCH_OK();
    return S2S_OK;
}

struct Plan {
    int terms, Kp;
    long rowsA, rowsB;          // rows per batch of A' / B'
    int nbA1, nbA2, nbB1, nbB2; // batch extents actually materialised (1 where the operand is broadcast)
    size_t bytesA, bytesB;
};

static Plan make_plan(const s2s_gemm_t& g) {
    Plan pl;
    pl.terms = g.split_terms == 3 ? 3 : 6;
    pl.Kp = (g.K + 7) / 8 * 8;
    pl.rowsA = (long)g.M + g.taps - 1;
    pl.rowsB = (long)g.N * g.taps;
    pl.nbA1 = (g.batch1 > 1 && g.a_bs1 != 0) ? g.batch1 : 1;
    pl.nbA2 = (g.batch2 > 1 && g.a_bs2 != 0) ? g.batch2 : 1;
    pl.nbB1 = (g.batch1 > 1 && g.b_bs1 != 0) ? g.batch1 : 1;
    pl.nbB2 = (g.batch2 > 1 && g.b_bs2 != 0) ? g.batch2 : 1;
    auto al = [](size_t n) { return (n + 255) / 256 * 256; };
    pl.bytesA = al((size_t)pl.nbA1 * pl.nbA2 * pl.rowsA * pl.terms * pl.Kp * sizeof(bf16));
    pl.bytesB = al((size_t)pl.nbB1 * pl.nbB2 * pl.rowsB * pl.terms * pl.Kp * sizeof(bf16));
    return pl;
}

}  // namespace split

size_t gemm_split_workspace_bytes(const s2s_gemm_t& g) {
    const split::Plan pl = split::make_plan(g);
    return pl.bytesA + pl.bytesB + 256;
}

int gemm_tc_split(const s2s_gemm_t& g, cudaStream_t st) {
    using namespace split;
    if (g.a_dtype != S2S_F32 || g.b_dtype != S2S_F32)
        return set_error(S2S_ERR_UNSUPPORTED, "gemm(mode 2): operands must be float32 (a=%d b=%d)", g.a_dtype, g.b_dtype);
    if (g.split_terms != 3 && g.split_terms != 6) return set_error(S2S_ERR_INVALID, "gemm(mode 2): split_terms must be 3 or 6");
    if (g.K <= 0) return set_error(S2S_ERR_INVALID, "gemm(mode 2): K must be positive");
    const Plan pl = make_plan(g);
    const size_t need = pl.bytesA + pl.bytesB + 256;
    if (!g.ws || g.ws_bytes < need) return set_error(S2S_ERR_INVALID, "gemm(mode 2): workspace of %zu bytes required (got %zu)", need, g.ws_bytes);
    uintptr_t base = (reinterpret_cast<uintptr_t>(g.ws) + 255) & ~(uintptr_t)255;
    bf16* Ap = reinterpret_cast<bf16*>(base);
    bf16* Bp = reinterpret_cast<bf16*>(base + pl.bytesA);
    const long ldk = (long)pl.terms * pl.Kp;
    Src sa{(const float*)g.A, g.a_rs, g.a_cs, pl.nbA1 > 1 ? g.a_bs1 : 0, pl.nbA2 > 1 ? g.a_bs2 : 0, (int)pl.rowsA, g.K, pl.Kp, pl.terms, 0, pl.nbA2};
    int rc = launch_split(sa, Ap, pl.nbA1 * pl.nbA2, st);
    if (rc != S2S_OK) return rc;
    // B rows are (n, tap) pairs: row index n * taps + t at element offset n * b_rs + t * b_ts
    if (g.taps > 1 && g.b_ts * g.taps != g.b_rs)
        return set_error(S2S_ERR_UNSUPPORTED, "gemm(mode 2): taps form needs B packed as (N, taps, K)");
    Src sb{(const float*)g.B, g.taps > 1 ? g.b_ts : g.b_rs, g.b_cs, pl.nbB1 > 1 ? g.b_bs1 : 0, pl.nbB2 > 1 ? g.b_bs2 : 0, (int)pl.rowsB, g.K, pl.Kp,
           pl.terms, 1, pl.nbB2};
    rc = launch_split(sb, Bp, pl.nbB1 * pl.nbB2, st);
    if (rc != S2S_OK) return rc;
    s2s_gemm_t h = g;
    h.A = Ap; h.a_dtype = S2S_BF16; h.a_rs = ldk; h.a_cs = 1;
    h.a_bs2 = pl.nbA2 > 1 ? pl.rowsA * ldk : 0;
    h.a_bs1 = pl.nbA1 > 1 ? (long)pl.nbA2 * pl.rowsA * ldk : 0;
    h.B = Bp; h.b_dtype = S2S_BF16; h.b_cs = 1;
    h.b_rs = (long)g.taps * ldk; h.b_ts = g.taps > 1 ? ldk : 0;
    h.b_bs2 = pl.nbB2 > 1 ? pl.rowsB * ldk : 0;
    h.b_bs1 = pl.nbB1 > 1 ? (long)pl.nbB2 * pl.rowsB * ldk : 0;
    h.K = (int)ldk;
    return gemm_tc(h, st);
}

}  // namespace s2s
