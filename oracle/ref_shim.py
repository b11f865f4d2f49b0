"""Import shim for the live reference (build container only; /root/reference is absent on GPU boxes).

TEST INFRASTRUCTURE ONLY.  The reference snapshot does not import cleanly:
 * seq2seq_vc/modules/alignments.py:226 eagerly compiles a numba function with an explicit
   signature that fails to type under numba 0.65 -> make that form of ``numba.jit`` lazy;
 * seq2seq_vc/losses/__init__.py:7 imports a module that is not in the tree -> stub it.
Nothing is copied: the reference is imported from where it lies.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("S2SVC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "seq2seq_vc"))


def install() -> None:
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import numba

    if not getattr(numba.jit, "_s2svc_lazy", False):
        orig = numba.jit

        def lazy_jit(*args, **kwargs):
            if args and not callable(args[0]):  # jit((signature...), nopython=True)
                return orig(**kwargs)
            return orig(*args, **kwargs)

        lazy_jit._s2svc_lazy = True
        numba.jit = lazy_jit
    sys.modules.setdefault("seq2seq_vc.losses.diffsinger_l2_loss",
                           types.ModuleType("seq2seq_vc.losses.diffsinger_l2_loss"))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def disable_dropout(model) -> None:
    """Zero every dropout incl. the always-on Prenet one (modules/pre_postnets.py:65)."""
    import torch

    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "dropout_rate"):
            m.dropout_rate = 0.0
