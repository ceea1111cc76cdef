python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline --no-kernel-pass 2>/dev/null | cut -c1-120
B200DIFF_ENCODER_STREAMS=0 python bench.py --no-cpu-baseline --no-kernel-pass 2>/dev/null | cut -c1-120
