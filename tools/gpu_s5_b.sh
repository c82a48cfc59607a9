# session-5 call B (2 GPUs): NCCL gather bit-equality, bench at N=2 (overlapped and sync gather), reference arm under torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/check_gather_nccl.py 64 2>&1 | grep -v Warning | tail -4
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b_bench_n2.json 2> gpurun_out/b_bench_n2.err; tail -3 gpurun_out/b_bench_n2.err
$TR bench.py --gpus 2 --steps 20 --warmup 5 --sync-gather --no-e2e > gpurun_out/b_bench_n2_sync.json 2>> gpurun_out/b_bench_n2.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench_n1.json 2>> gpurun_out/b_bench_n2.err
python - <<'PY'
import json
for f in ("b_bench_n1", "b_bench_n2", "b_bench_n2_sync"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "n", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", d.get("e2e", {}).get("ms_per_step"), d["config"]["gather"])
    except Exception as e:
        print(f, "failed", e)
PY
