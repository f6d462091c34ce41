timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_context_bank_gpu.py -m gpu -x -q -k "attention or context or full_window or reference_style" 2>&1 | tail -15 > gpurun_out/pytest_ltc3.log
tail -c 1200 gpurun_out/pytest_ltc3.log
TUBER_KPROF_DUMP=1 timeout 200 python tools/run_ltc.py > gpurun_out/ltc_v33.json 2> gpurun_out/ltc_kprof_v33.txt
cat gpurun_out/ltc_v33.json | cut -c1-500; grep "attention" gpurun_out/ltc_kprof_v33.txt | tail -3
TUBER_ATTN_NO_PREP=1 TUBER_KPROF_DUMP=1 timeout 200 python tools/run_ltc.py 2>&1 | grep "L=1024 S=16384"
