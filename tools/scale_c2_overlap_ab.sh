# usage (under gpurun --gpus N): bash tools/scale_c2_overlap_ab.sh N -> gpurun_out/r02_scale_c2_nN_overlap{0,1}.json
N=$1
for OV in 0 1; do
  S2S_OVERLAP_ALLREDUCE=$OV python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_c2_n${N}_overlap$OV.json 2> gpurun_out/r02_scale_c2_n${N}_overlap$OV.err
  python -c "
import json
for l in open('gpurun_out/r02_scale_c2_n${N}_overlap$OV.json'):
    if l.startswith('{'):
        d=json.loads(l); print('overlap=$OV', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'])
"
done
