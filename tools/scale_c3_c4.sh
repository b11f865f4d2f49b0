# usage (under gpurun --gpus N): bash tools/scale_c3_c4.sh N  -> gpurun_out/r02_scale_{c3,c4,c2}_nN.json
N=$1
for W in c3 c4 c2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_${W}_n$N.json 2> gpurun_out/r02_scale_${W}_n$N.err
  tail -c 400 gpurun_out/r02_scale_${W}_n$N.json | head -c 400; echo
done
