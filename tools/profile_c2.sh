set -x
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2.csv python tools/profile_step.py c2 2>&1 | tail -1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_probs_fwd_kernel -s 8 -c 1 -o gpurun_out/c2_attn_probs_fwd_full -f python tools/profile_step.py c2 2>&1 | tail -1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_tc_metrics_c2.csv python tools/profile_step.py c2 2>&1 | tail -1
timeout 300 python bench.py --gemm-table gpurun_out/gemm_c2_final.txt 2>&1 | tail -1 > gpurun_out/bench_c2_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c2_reference.json
timeout 400 python bench.py --workload c3 --gemm-table gpurun_out/gemm_c3_final.txt 2>&1 | tail -1 > gpurun_out/bench_c3_final.json
timeout 200 python tools/micro_bench.py 2>&1 | tail -12
