python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; tail -3 gpurun_out/q_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/q_bench.json'))
print("ms/step",d["ms_per_step"],"e2e",d.get("e2e",{}).get("ms_per_step"))
for k,v in d["kernels"].items(): print(k, round(v["ms"],3))
print(d["contact"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python tools/launches.py gpurun_out/q_launches.csv 8
