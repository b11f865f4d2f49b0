# Round-2 closing evidence after the last C3-side kernel changes (tensor-core alignment distances, one-MUFU sigmoid): GPU test log on the
# final tree, the AAS-VC / FastSpeechVC bench lines and launch lists again (run under gpurun on one B200).
set -x
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_final.log
for W in c3 c3s c6; do
  python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/r02_final_bench_$W.json 2> gpurun_out/r02_final_bench_$W.err
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_c3.csv python tools/profile_step.py c3 2>&1 | tail -1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_final_launches_c6.csv python tools/profile_step.py c6 2>&1 | tail -1
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_final_bench_c[36]*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f.split('bench_')[1], round(d.get('ms_per_step',0),3), round(d.get('value',0)), d.get('roofline',{}).get('frac'), d.get('gpu_launches'), d.get('clocks'))
"
