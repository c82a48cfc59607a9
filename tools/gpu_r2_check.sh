# round-2 check of the working tree: GPU tests, smoke, short bench (default + falloff none)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $BENCH_EXTRA > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -15 gpurun_out/c_pytest.log; tail -2 gpurun_out/c_smoke.log; tail -3 gpurun_out/c_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/c_bench.json'))
    print("ms/step",round(d["ms_per_step"],3),"e2e",round(d.get("e2e",{}).get("ms_per_step",0),3), "launches", d["gpu_launches"], {k:round(v["ms"],3) for k,v in d["kernels"].items()})
    print("roofline", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","ms_per_launch")})
except Exception as ex:
    print("bench failed", ex)
PY
