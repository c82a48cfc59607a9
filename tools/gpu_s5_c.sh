# session-5 call C (2 GPUs): NCCL gather bit-equality, then 1-GPU overlap A/B + fill-split sweep
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/check_gather_nccl.py 64 > gpurun_out/c_gather.log 2>&1; grep -v Warn gpurun_out/c_gather.log | grep -i "nccl gather\|Error\|assert" | head -5
$TR bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e 2> gpurun_out/c_n2.err | head -c 300; echo
for v in "" "--no-overlap"; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline $v > gpurun_out/c_bench$v.json 2> gpurun_out/c_bench$v.err; tail -3 gpurun_out/c_bench$v.err
  python - "$v" <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/c_bench{sys.argv[1]}.json'))
print(sys.argv[1] or "overlap", "ms/step",round(d["ms_per_step"],3),"e2e",d.get("e2e",{}).get("ms_per_step"), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
print(d["clocks"], d["roofline"]["frac"])
PY
done
