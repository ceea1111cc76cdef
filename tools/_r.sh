python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for f in "" s r srb; do python tools/conv_one.py 8 256 256 320 320 3 "$f"; done
python tools/conv_one.py 8 256 256 96 96 3 "srb"
python tools/conv_one.py 8 32 32 640 1920 1 ""
python bench.py --no-cpu-baseline > gpurun_out/bench11.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench11.json"))
print(d["value"], d["e2e"]["value"], d["clocks"], d["roofline"]["achieved"], d["kernel_pass_ms_per_denoise_step"])
for a in d["roofline_aux"][:6]: print(a["kernel"], round(a["achieved"]), round(a["ms_per_denoise_step"],2))
PY
