"""Micro-benchmark of s2s_gemm(mode=1) shapes (CUDA events, L2 flushed between launches)."""
import sys, os, math, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seq2seq_vc_b200 import ops

SHAPES = {
    "lin_dec": dict(M=16384, N=384, K=384), "ffn1": dict(M=16384, N=1536, K=384), "ffn2": dict(M=16384, N=384, K=1536),
    "lin_enc": dict(M=4064, N=384, K=384), "conv2": dict(M=77216, N=384, K=3456), "big": dict(M=8192, N=8192, K=8192),
    "qk": dict(M=512, N=512, K=48, batch=256), "c3_ffn": dict(M=49152, N=1536, K=1536), "c3_qkv": dict(M=49152, N=4608, K=1536),
    "c3_qk": dict(M=768, N=768, K=768, batch=128), "c3_ffn_in": dict(M=49152, N=1536, K=384), "c3_ffn_out": dict(M=49152, N=384, K=1536),
    "c3_lin": dict(M=49152, N=384, K=384), "pv": dict(M=512, N=48, K=512, batch=256),
    "c3_ffn3": dict(M=49152, N=3072, K=1536), "c3_ffn3o": dict(M=49152, N=1536, K=3072), "c3_qkvo": dict(M=49152, N=1536, K=4608),
    "c2b64_lin": dict(M=32768, N=384, K=384), "c2b64_ffn1": dict(M=32768, N=1536, K=384), "c4_lin": dict(M=32000, N=512, K=512), "one_tile": dict(M=128, N=128, K=384), "one_tile_longk": dict(M=128, N=128, K=8192),
}

def _setup(name):
    s = SHAPES[name]
    nb = s.get("batch", 1)
    M, N, K = s["M"], s["N"], s["K"]
    a = torch.randn(nb, M, K, device="cuda").bfloat16()
    b = torch.randn(nb, N, K, device="cuda").bfloat16()
    c = torch.empty(nb, M, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    if nb == 1:
        a, b, c = a[0], b[0], c[0]
    else:
        a, b, c = a.view(nb // 8, 8, M, K), b.view(nb // 8, 8, N, K), c.view(nb // 8, 8, M, N)
    return a, b, c, (bias if nb == 1 else None), 2.0 * nb * M * N * K, f"{nb}x{M}x{N}x{K}"


_scratch = None


FULL = False      # residual + dropout epilogue (EPI_BF16_FULL) instead of bias only
NOBIAS = False
_res = {}


def _time_once(a, b, c, bias, flush=True):
    global _scratch
    kw = {}
    if FULL:
        from seq2seq_vc_b200._lib import Drop
        if c.data_ptr() not in _res:
            _res.clear()
            _res[c.data_ptr()] = torch.randn_like(c)
        kw = {}
        if FULL in (True, "res"):
            kw["residual"] = _res[c.data_ptr()]
        if FULL in (True, "drop"):
            kw["drop"] = Drop(0.1, seed=3, site=2)
        if FULL == "gate":
            kw["gate"] = _res[c.data_ptr()]
    if _scratch is None:
        _scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if flush:
        _scratch.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(a, b, c, bias=None if NOBIAS else bias, mode=1, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


def run(name, reps=10, flush=True):
    a, b, c, bias, fl, desc = _setup(name)
    ts = [_time_once(a, b, c, bias, flush) for _ in range(reps + 3)][3:]
    ts.sort()
    us = ts[len(ts) // 2]
    print(f"{name:16s} {desc}: median {us:8.1f} us  min {ts[0]:8.1f} us  {fl / us / 1e6:8.1f} TFLOP/s", flush=True)


def ab(names, rounds=9):
    """Single-CTA tiles vs CTA pairs (s2s_debug_gemm_tile) vs the cost model's own choice, INTERLEAVED launch by launch:
    the box's power / clock state drifts by more than the difference being measured."""
    from seq2seq_vc_b200 import _lib
    lib = _lib.load()
    for n in names:
        a, b, c, bias, fl, desc = _setup(n)
        ts = {1: [], 2: [], 0: []}
        for r in range(rounds + 2):
            for cg in (1, 2, 0):
                lib.s2s_debug_gemm_tile(cg)
                t = _time_once(a, b, c, bias)
                if r >= 2:
                    ts[cg].append(t)
        med = {cg: sorted(v)[len(v) // 2] for cg, v in ts.items()}
        print(f"{n:12s} {desc:22s} single {med[1]:8.1f} us {fl / med[1] / 1e6:7.1f} TF | pair {med[2]:8.1f} us {fl / med[2] / 1e6:7.1f} TF"
              f" | model {med[0]:8.1f} us {fl / med[0] / 1e6:7.1f} TF   pair/single {med[1] / med[2]:.3f}", flush=True)
    lib.s2s_debug_gemm_tile(0)


def bn_sweep(names, bns=(64, 96, 128, 192, 256), rounds=7):
    """Pinned N-tile widths vs the cost model's choice (0), interleaved launch by launch, cold (L2 flushed) and warm operands."""
    from seq2seq_vc_b200 import _lib
    lib = _lib.load()
    for n in names:
        a, b, c, bias, fl, desc = _setup(n)
        for flush in (True, False):
            ts = {bn: [] for bn in (0,) + tuple(bns)}
            for r in range(rounds + 2):
                for bn in ts:
                    lib.s2s_debug_gemm_tile(0)
                    lib.s2s_debug_gemm_tile(1)
                    if bn:
                        lib.s2s_debug_gemm_tile(bn)
                    else:
                        lib.s2s_debug_gemm_tile(0)
                    t = _time_once(a, b, c, bias, flush)
                    if r >= 2:
                        ts[bn].append(t)
            med = {bn: sorted(v)[len(v) // 2] for bn, v in ts.items()}
            if flush:      # pinned tiles must give the model-chosen tile's result (same k order per output: bit-equal)
                lib.s2s_debug_gemm_tile(0)
                ops.gemm(a, b, c, bias=bias, mode=1)
                ref = c.clone()
                for bn in bns:
                    lib.s2s_debug_gemm_tile(1)
                    lib.s2s_debug_gemm_tile(bn)
                    c.zero_()
                    ops.gemm(a, b, c, bias=bias, mode=1)
                    assert torch.equal(ref, c), f"{n}: BN {bn} differs from the model tile by {(ref.float() - c.float()).abs().max().item()}"
                lib.s2s_debug_gemm_tile(0)
            print(f"{n:12s} {desc:20s} {'cold' if flush else 'warm'} " + " | ".join(f"{'model' if bn == 0 else 'BN' + str(bn)} {med[bn]:6.1f} us" for bn in ts),
                  flush=True)
    lib.s2s_debug_gemm_tile(0)


def trace(name, flush=True):
    import ctypes
    from seq2seq_vc_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    lib.s2s_debug_gemm_trace.argtypes = [ctypes.c_void_p]
    lib.s2s_debug_gemm_trace(buf.data_ptr())
    run(name, reps=2, flush=flush)
    t = buf.cpu().tolist()
    lib.s2s_debug_gemm_trace(None)
    t0 = t[0]
    print(name, "flush" if flush else "warm", "cycles: setup %d, first TMA issue %d, kernel end %d" % (t[1] - t0, t[2] - t0, t[7] - t0))
    launches = 5   # run(reps=2) = 2 + 3 warm-up launches accumulate into the wait counters
    print("   CTA 0 per launch: %d k-blocks; MMA warp blocked on full %d cyc, on tmem_empty %d cyc; TMA producer blocked on empty %d cyc"
          % (t[11] // launches, t[8] // launches, t[9] // launches, t[10] // launches))
    for i in range(4):
        r = t[16 + 8 * i: 16 + 8 * i + 8]
        if r[0] == 0:
            break
        print("   tile %d: MMA first-full %d last-full %d | epilogue warp 2: tfull %d, chunk 0 in registers %d, computed %d, tmem released %d, done %d" %
              (i, r[0] - t0, r[1] - t0, r[2] - t0, r[6] - t0, r[7] - t0, r[3] - t0, r[4] - t0))


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] in ("full", "res", "drop", "gate"):
        FULL = True if args[0] == "full" else args[0]
        args = args[1:]
    if args and args[0] == "nobias":
        NOBIAS = True
        args = args[1:]
    if args and args[0] == "ab":
        ab(args[1:] or list(SHAPES))
        sys.exit(0)
    if args and args[0] == "bn":
        bn_sweep(args[1:] or ["lin_dec", "ffn1", "lin_enc", "c2b64_lin"])
        sys.exit(0)
    if args and args[0] == "trace":      # needs a library built with -DS2S_GEMM_TRACE
        from seq2seq_vc_b200 import _lib
        for n in args[1:]:
            for cg in (1,):
                _lib.load().s2s_debug_gemm_tile(cg)
                print("cg=%d" % cg)
                trace(n, True)
    else:
        for n in (args or list(SHAPES)):
            run(n)
