# session-5 call A: full GPU parity suite + smoke, per-phase cycle shares (CT_PROFILE build), kernel table
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
IGI_NVCC_EXTRA="-DCT_PROFILE" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
python tools/ct_prof.py 2>&1 | tee gpurun_out/a_ctprof.txt
python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -3 gpurun_out/a_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/a_bench.json'))
print("ms/step",round(d["ms_per_step"],3),"e2e",d.get("e2e",{}).get("ms_per_step"), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
print(d["clocks"], d["contact"])
PY
