set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -5
for mode in "TUBER_NO_OVERLAP=1" "TUBER_SIDE_CTAS=0" "TUBER_SIDE_CTAS=64" "TUBER_SIDE_CTAS=96"; do
  env $mode python bench.py --no-also --no-cpu-baseline --steps 100 > gpurun_out/ov_$mode.json 2> gpurun_out/ov_$mode.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/ov_$mode.json").read().strip().splitlines()[-1])
print("$mode", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["e2e"]["value"])
PY
done
