# usage: gpu_r2_e2e_gather.sh N : e2e leg under both gather transports, alternating (copy-engine contention check)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for g in p2p nccl p2p nccl; do
  timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --gather $g --no-components --no-alt-falloff > gpurun_out/eg_$g.json 2> gpurun_out/eg_$g.err
  python - $g <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/eg_{sys.argv[1]}.json"))
print(sys.argv[1], "resident ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3))
PY
done
