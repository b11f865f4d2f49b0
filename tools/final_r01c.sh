# A/B of programmatic dependent launch for the single-CTA tcgen05 GEMMs (S2S_GEMM_PDL=1): parity first, then bench lines.
mkdir -p gpurun_out
S2S_GEMM_PDL=1 timeout 60 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vtn.py -x -q -m gpu > gpurun_out/pdl_pytest_a.log 2>&1; echo "pdl tests (kernels+vtn) rc=$?"
S2S_GEMM_PDL=1 timeout 50 python bench.py --no-cpu-baseline 2> gpurun_out/pdl_bench_c2_on.err | tail -1 > gpurun_out/pdl_bench_c2_on.json; echo "c2 pdl on rc=$?"
timeout 50 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/pdl_bench_c2_off.json; echo "c2 pdl off rc=$?"
S2S_GEMM_PDL=1 timeout 50 python bench.py --workload c4 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/pdl_bench_c4_on.json; echo "c4 pdl on rc=$?"
S2S_GEMM_PDL=1 timeout 60 python bench.py --workload c3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/pdl_bench_c3_on.json; echo "c3 pdl on rc=$?"
S2S_GEMM_PDL=1 timeout 60 python -m pytest tests/test_gpu_aasvc.py tests/test_gpu_mas_logmel.py -x -q -m gpu > gpurun_out/pdl_pytest_b.log 2>&1; echo "pdl tests (aasvc) rc=$?"
for f in gpurun_out/pdl_bench_*.json; do echo $f; cut -c1-330 $f | grep -o '"ms_per_step": [0-9.]*'; done
tail -n 2 gpurun_out/pdl_pytest_a.log gpurun_out/pdl_pytest_b.log
