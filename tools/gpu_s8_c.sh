# session-8 call C: all GPU tests, bench, one ncu --set full capture of tac_contact (source lines kept)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; tail -3 gpurun_out/c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c_bench.json'))
print("ms/step",round(d["ms_per_step"],3),"e2e",d.get("e2e",{}).get("ms_per_step"), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
bash tools/gpu_ncu1.sh tac_contact contact_s8
