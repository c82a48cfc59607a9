# usage: gpu_variants2.sh "<flags A>" "<flags B>" ...   -> tactile parity + bench kernel table per build variant
for v in "$@"; do
  echo "=== variant: $v"
  IGI_NVCC_EXTRA="$v" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  python -m pytest tests/test_tactile_gpu.py tests/test_task_gpu.py -m gpu -x -q 2>&1 | tail -1
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -3 gpurun_out/v_bench.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/v_bench.json'))
print("ms/step",round(d["ms_per_step"],3), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
done
