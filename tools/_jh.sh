python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "sgemm" 2>&1 | tail -3
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "golden or real_evaluation" 2>&1 | tail -3
TUBER_KPROF_DUMP=1 python bench.py --config Tuber_CSN152_JHMDB.yaml --clip 16 256 256 --no-also --no-cpu-baseline --steps 100 > gpurun_out/jh2.json 2> gpurun_out/jh2.err
grep -E "kprof (252|253|256|272) " gpurun_out/jh2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/jh2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["stage_ms"])
PY
