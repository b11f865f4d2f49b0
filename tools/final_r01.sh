# End-of-round evidence run (one gpurun call): GEMM parity first, then the C2 bench line + per-shape GEMM table,
# the ncu launch list of one C2 step, the C3 bench line, then the whole GPU suite with whatever time is left.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm" > gpurun_out/final_pytest_gemm.log 2>&1; echo "gemm tests rc=$?"
timeout 240 python bench.py --gemm-table gpurun_out/final_gemm_c2.txt 2> gpurun_out/final_bench_c2.err | tail -1 > gpurun_out/final_bench_c2.json; echo "bench c2 rc=$?"
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_c2.csv python tools/profile_step.py c2 2>&1 | tail -1
timeout 240 python bench.py --workload c3 --no-cpu-baseline --gemm-table gpurun_out/final_gemm_c3.txt 2> gpurun_out/final_bench_c3.err | tail -1 > gpurun_out/final_bench_c3.json; echo "bench c3 rc=$?"
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_c3.csv python tools/profile_step.py c3 2>&1 | tail -1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -3 gpurun_out/final_pytest_gemm.log gpurun_out/final_pytest_gpu.log
cat gpurun_out/final_bench_c2.json | cut -c1-400
