python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
python tools/step_breakdown.py cfg3 8 > gpurun_out/breakdown_cfg3.txt 2>&1; head -24 gpurun_out/breakdown_cfg3.txt
