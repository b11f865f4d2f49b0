"""CPU oracle for the seq2seq-vc hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(plain PyTorch fp32 for the floating-point path, numpy / C for the integer
alignment path).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker or as the CPU baseline -- never as the thing shipped.
``seq2seq_vc_b200`` must not import this package.
"""
