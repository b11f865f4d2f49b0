set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python tools/profile_step.py c3 2>&1 | tail -1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none -k regex:gemm_tc_kernel --csv --log-file gpurun_out/gemm_tc_metrics_c3.csv python tools/profile_step.py c3 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dwconv_tile_kernel -c 1 -o gpurun_out/c3_dwconv_full -f python tools/micro_aas.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 1 -o gpurun_out/c3_gemm_ffn_full -f python tools/gemm_probe.py c3_ffn 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_reg_kernel -c 1 -o gpurun_out/c3_ln_bwd_full -f python tools/micro_aas.py 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pairdist_kernel -c 1 -o gpurun_out/c3_pairdist_full -f python tools/micro_aas.py 2>&1 | tail -1
timeout 200 python tools/gemm_probe.py c3_ffn c3_qkv c3_qk 2>&1 | tail -3
timeout 200 python tools/micro_aas.py gpurun_out/micro_aas_final.json 2>&1 | tail -32
