# Last check of the final tree (default environment): smoke(), the whole GPU suite, the default bench line.
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
timeout 60 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
timeout 100 python bench.py 2> gpurun_out/final_bench_c2.err | tail -1 > gpurun_out/final_bench_c2.json; echo "bench rc=$?"
tail -n 2 gpurun_out/final_smoke.log gpurun_out/final_pytest_gpu.log; cut -c1-300 gpurun_out/final_bench_c2.json
