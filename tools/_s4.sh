timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "test_stem" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_frames_gpu.py -m gpu -x -q 2>&1 | tail -3
for mode in "TUBER_STEM_NO_VEC=1" "TUBER_STEM_NO_VEC=0"; do
  env $mode python bench.py --no-also --no-cpu-baseline --steps 100 > gpurun_out/s4_$mode.json 2> gpurun_out/s4_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s4_$mode.json").read().strip().splitlines()[-1])
print("$mode", round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], d["stage_ms"]["stem"], [k["ms"] for k in d["kernels"] if k["kernel"]=="stem_conv"])
PY
done
