# Final evidence of round 2 (run under gpurun on one B200): GPU test log, one bench line per workload, launch lists.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_c2.json 2> gpurun_out/r02_final_bench_c2.err
for W in c2b64 c4 c1 c3 c3s c5 c6; do
  python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/r02_final_bench_$W.json 2> gpurun_out/r02_final_bench_$W.err
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference_c2.json 2> gpurun_out/r02_final_bench_reference_c2.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_c2.csv python tools/profile_step.py c2 2>&1 | tail -1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_c3.csv python tools/profile_step.py c3 2>&1 | tail -1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_c6.csv python tools/profile_step.py c6 2>&1 | tail -1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ln_bwd_grad_kernel -s 20 -c 1 -o gpurun_out/r02_c2_ln_bwd_grad_full -f python tools/profile_step.py c2 2>&1 | tail -1
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_bench_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f.split('bench_')[1], round(d.get('ms_per_step',0),3), round(d.get('value',0)), d.get('roofline',{}).get('frac'), d.get('gpu_launches'), d.get('clocks'))
"
