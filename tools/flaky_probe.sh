# Order-dependence probe for the CUDA-graph tests (run under gpurun): the sequence that used to invalidate a capture, three times,
# then the whole GPU suite twice.
r() { echo "== $1"; shift; "$@" 2>&1 | grep -E "^(FAILED|[0-9]+ (passed|failed))|invalidated" | cut -c1-300 | head -8; }
for i in 1 2 3; do r "kernels+vtn #$i" python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vtn.py -q; done
r "kernels(layernorm)+vtn(graph tests)" python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vtn.py -q -k "layernorm or alternating or eviction"
for i in 1 2; do r "full suite #$i" python -m pytest tests -m gpu -q; done
