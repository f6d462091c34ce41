for dbg in 0 1 2 4 8 3 15; do
  TUBER_STEM3=1 TUBER_STEM_DBG=$dbg python bench.py --no-also --no-cpu-baseline --steps 20 > gpurun_out/s3d_$dbg.json 2> gpurun_out/s3d_$dbg.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/s3d_$dbg.json").read().strip().splitlines()[-1])
print("dbg=$dbg", round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], "stem stage", d["stage_ms"]["stem"], "kprof", [k["ms"] for k in d["kernels"] if k["kernel"]=="stem_conv"])
PY
done
