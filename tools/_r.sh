timeout 300 python tools/attn_check.py 2>&1 | grep -E "float16"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
