# usage: gpu_variants.sh "<flags A>" "<flags B>" ...   -> bench kernel table per variant
for v in "$@"; do
  echo "=== variant: $v"
  IGI_NVCC_EXTRA="$v" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  python -m pytest tests/test_tactile_gpu.py -m gpu -x -q 2>&1 | tail -2
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -3 gpurun_out/v_bench.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/v_bench.json'))
print("ms/step",round(d["ms_per_step"],3), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
print(d["contact"])
PY
done
