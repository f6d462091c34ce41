timeout 200 python -m pytest tests/test_context_bank_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_ltc2.log
tail -c 1200 gpurun_out/pytest_ltc2.log
TUBER_KPROF_DUMP=1 timeout 200 python tools/run_ltc.py > gpurun_out/ltc_v32.json 2> gpurun_out/ltc_kprof_v32.txt
cat gpurun_out/ltc_v32.json | cut -c1-600; grep "attention" gpurun_out/ltc_kprof_v32.txt | tail -4
TUBER_ATTN_NO_TC=1 TUBER_KPROF_DUMP=1 timeout 200 python tools/run_ltc.py 2>&1 | grep "L=1024 S=16384"
