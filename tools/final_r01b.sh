# Second end-of-round evidence run: the gradient-accumulation GPU test + whole GPU suite, refreshed ncu captures of the
# tcgen05 GEMM (one full capture at the C3 FFN shape, tensor-pipe / DRAM metrics over every GEMM launch of a C2 and a C3
# step), and the bench lines of the remaining workloads.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 120 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 1 -o gpurun_out/c3_gemm_ffn_full -f python tools/gemm_probe.py c3_ffn 2>&1 | tail -1
timeout 60 python tools/ncu_brief.py gpurun_out/c3_gemm_ffn_full.ncu-rep > gpurun_out/c3_gemm_ffn_full_brief.txt 2>&1
timeout 60 python tools/gemm_probe.py c3_ffn c3_qkv c3_qk big 2>&1 | tail -4 > gpurun_out/final_gemm_probe.txt
timeout 100 ncu --profile-from-start off --metrics $M --clock-control none -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_tc_metrics_c2.csv python tools/profile_step.py c2 2>&1 | tail -1
timeout 150 ncu --profile-from-start off --metrics $M --clock-control none -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_tc_metrics_c3.csv python tools/profile_step.py c3 2>&1 | tail -1
timeout 90 python bench.py --workload c2b64 --no-cpu-baseline 2> gpurun_out/final_bench_c2b64.err | tail -1 > gpurun_out/final_bench_c2b64.json; echo "bench c2b64 rc=$?"
timeout 90 python bench.py --workload c4 --no-cpu-baseline 2> gpurun_out/final_bench_c4.err | tail -1 > gpurun_out/final_bench_c4.json; echo "bench c4 rc=$?"
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final_bench_c2_reference.json; echo "reference arm rc=$?"
tail -n 3 gpurun_out/final_pytest_gpu.log
cat gpurun_out/final_gemm_probe.txt
