run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps $K --warmup 5 --no-e2e 2>gpurun_out/ab_$name.err | tail -1 > gpurun_out/ab_$name.json
  python -c "
import json;d=json.load(open('gpurun_out/ab_$name.json'));print('$name', $K, round(d['value'],1), round(d['ms_per_step'],4))" || tail -5 gpurun_out/ab_$name.err
}
for K in 20 100; do
run peer_fence_$K OADG_EXCHANGE=peer OADG_CONSUMER_FENCE=1
run nccl_fence_$K OADG_EXCHANGE=nccl OADG_CONSUMER_FENCE=1
run peer_nofence_$K OADG_EXCHANGE=peer OADG_CONSUMER_FENCE=0
done
