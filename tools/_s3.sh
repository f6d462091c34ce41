TUBER_STEM3=1 timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "golden or stages or real_evaluation" 2>&1 | tail -3
for mode in "TUBER_STEM3=0" "TUBER_STEM3=1"; do
  env $mode python bench.py --no-also --no-cpu-baseline --steps 100 > gpurun_out/s3_$mode.json 2> gpurun_out/s3_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s3_$mode.json").read().strip().splitlines()[-1])
print("$mode", round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], d["stage_ms"]["stem"], [k for k in d["kernels"] if k["kernel"]=="stem_conv"])
PY
done
