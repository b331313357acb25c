timeout 600 python -m pytest tests/test_ops_f16_gpu.py tests/test_model_gpu.py -q -x 2>&1 | grep -E "Error|assert|FAILED|rel\(" | head -12
