# round-1 session-6 call A: parity suite with the warp-cooperative pcl_compact, phase shares of tac_contact, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/a_pytest.log; cat gpurun_out/a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -3 gpurun_out/a_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/a_bench.json'))
print("ms/step",round(d["ms_per_step"],3),"e2e",d.get("e2e",{}).get("ms_per_step"), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
IGI_NVCC_EXTRA="-DCT_PROFILE" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
python tools/ct_prof.py > gpurun_out/a_ctprof.txt 2>&1; cat gpurun_out/a_ctprof.txt
python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
