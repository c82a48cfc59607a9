# session-8 call D: tactile/task parity after a contact change, bench without the CPU legs, optional ncu capture ($1 = kernel regex)
mkdir -p gpurun_out
python -m pytest tests/test_tactile_gpu.py tests/test_task_gpu.py -m gpu -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; tail -3 gpurun_out/d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d_bench.json'))
print("ms/step",round(d["ms_per_step"],3), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
if [ -n "$1" ]; then bash tools/gpu_ncu1.sh $1 $2; fi
