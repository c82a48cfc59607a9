# e2e leg: uint8 vs int32 host segmentation, plus the host-pipeline test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_task_gpu.py -m gpu -q -x -k "host_pipeline or graphed" 2>&1 | tail -3
for extra in "" "--seg-int32-host" "" "--seg-int32-host"; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-components --no-alt-falloff $extra > gpurun_out/e2e_v.json 2> gpurun_out/e2e_v.err
  python - "$extra" <<'PY'
import json, sys
d = json.load(open("gpurun_out/e2e_v.json"))
e = d["e2e"]
print(repr(sys.argv[1]), "resident ms", round(d["ms_per_step"], 3), "e2e ms", round(e["ms_per_step"], 3), "h2d MB", round(e["h2d_bytes_per_step"] / 1e6, 1), e.get("seg_host_dtype"))
PY
done
