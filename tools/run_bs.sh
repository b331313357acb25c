timeout 600 python -m pytest tests/test_model_gpu.py -q -s -k bench_size 2>&1 | grep -E "rel err|passed|failed|Error" | head
