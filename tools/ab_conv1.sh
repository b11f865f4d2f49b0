python -m pytest tests/test_gpu_kernels.py -q -k "conv1_weight_gradient or conv2d" 2>&1 | tail -3
for f in 0 1 0 1; do S2S_CONV1_FWD_TC=$f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-incumbent 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CONV1_FWD_TC=$f', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['losses_last_step'])"; done
python -m pytest tests/test_gpu_vtn.py tests/test_gpu_fsvc.py tests/test_gpu_aasvc.py -q 2>&1 | tail -3
