# builds tactile variants ON the box (IGI_NVCC_EXTRA) and prints the per-kernel bench times of each
mkdir -p gpurun_out
for v in "$@"; do
  IGI_NVCC_EXTRA="$v" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -2 gpurun_out/v_bench.err
  python - "$v" <<'PY'
import json, sys
d=json.load(open('gpurun_out/v_bench.json'))
print(repr(sys.argv[1]), "ms/step",round(d["ms_per_step"],3), {k:round(v["ms"],3) for k,v in d["kernels"].items() if k.startswith("tac")})
PY
done
