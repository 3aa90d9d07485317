timeout 200 python scripts/conv_micro.py out3 64 10 2>&1 | tail -2
timeout 900 python -m pytest tests/test_vunet_gpu.py tests/test_pipeline_gpu.py tests/test_icn_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python scripts/step_time.py 30 2>&1 | tail -2
