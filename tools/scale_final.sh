# usage (under gpurun --gpus N): bash tools/scale_final.sh N  -> gpurun_out/r02_final_scale_{c2,c3,c6}_nN.json (final tree)
N=$1
for W in c2 c3 c6; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_scale_${W}_n$N.json 2> gpurun_out/r02_final_scale_${W}_n$N.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r02_final_scale_${W}_n$N.json').read().strip().splitlines()[-1])
print('$W', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['clocks'])"
done
