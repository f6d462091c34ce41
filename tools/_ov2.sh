python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "side_branches" 2>&1 | tail -3
for mode in "TUBER_X=0" "TUBER_FUSE2_L3=1" "TUBER_FUSE2_L3=1 TUBER_FUSE2_SINGLE=1" "TUBER_OVERLAP=1"; do
  for b in 2 4; do
  env $mode python bench.py --no-also --no-cpu-baseline --steps 150 --batch $b > "gpurun_out/b${b}_$mode.json" 2> "gpurun_out/b${b}_$mode.err"
  python - <<PY
import json
d=json.loads(open("gpurun_out/b${b}_$mode.json").read().strip().splitlines()[-1])
print("$mode", "batch", $b, round(d["value"],1), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], d["stage_ms"])
PY
  done
done
