python bench.py --no-cpu-baseline > gpurun_out/bench13.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench13.json"))
print(d["value"], d["e2e"]["value"], d["clocks"], d["roofline"]["achieved"], d["kernel_pass_ms_per_denoise_step"])
for a in d["roofline_aux"][:6]: print(a["kernel"], round(a["achieved"]), round(a["ms_per_denoise_step"],2))
PY
python tools/step_breakdown.py cfg3 8 2>&1 | head -4
