# usage: gpu_ctprof.sh [extra nvcc flags]: tactile parity tests, per-phase cycle shares, stage timings
IGI_NVCC_EXTRA="-DCT_PROFILE $1" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
python -m pytest tests/test_tactile_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/ct_prof.py
IGI_NVCC_EXTRA="$1" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -3 gpurun_out/v_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/v_bench.json'))
print("ms/step",round(d["ms_per_step"],3), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
