python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/step_breakdown.py cfg1 8 2>&1 | grep -E "cfg1|c8 "
