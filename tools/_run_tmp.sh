timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" 2>&1 | tail -3
timeout 300 python tools/gemm_probe.py ab lin_dec ffn1 ffn2 lin_enc conv2 qk pv c3_ffn c3_qkv c3_ffn_in c3_lin 2>&1 | tail -12
