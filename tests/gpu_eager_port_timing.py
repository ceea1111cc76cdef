"""Measurement helper (not a pytest module): the oracle port of the reference network run EAGERLY ON THE GPU, i.e. what the
reference itself costs on a B200 with stock PyTorch kernels (cuDNN / cuBLAS), per SURVEY.md §8(d) "also report the GPU-eager
reference (fp32 TF32-on and bf16-autocast)".  The oracle is the torch-functional restatement of the reference modules, so this
is the reference's kernel mix (conv2d, group_norm, silu, softmax attention with the L x L matrix materialised) without the
Lightning wrapper.  Lives under tests/ because only tests may execute oracle/.

    python tests/gpu_eager_port_timing.py --workload cfg2 --batch 8 > gpurun_out/eager.jsonl
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import bench  # noqa: E402  (workload table + synthetic inputs only)
import dsdiff_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    wl = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda:0")
    sd = {k: v.to(dev) for k, v in O.synth_state_dict(bench.meta_shapes(wl)).items()}
    cond, x_T = bench.synth_inputs(a.batch, wl["cond_ch"], a.size, 0)
    xc = torch.cat([x_T, cond], 1).to(dev)
    ts = torch.full((a.batch,), 1961, dtype=torch.long, device=dev)
    fwd = O.dsunet_forward if wl["cls"] == "DSUnetModel" else O.unet_forward
    ref = None
    for mode in ("fp32_strict", "fp32_tf32", "bf16_autocast"):
        tf32 = mode != "fp32_strict"
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16_autocast"):
                out = fwd(sd, wl["hp"], xc, ts)
            return (out[0] if isinstance(out, tuple) else out).float()

        for _ in range(2):
            out = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            out = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        if ref is None:
            ref = out
        print(json.dumps({
            "what": "oracle port of the reference network, eager PyTorch on cuda:0 (network forward only, no sampler update)",
            "workload": a.workload, "mode": mode, "batch": a.batch, "ms_per_forward": round(ms, 2),
            "slices_per_s_ddim50": round(a.batch / (50 * ms * 1e-3), 4),
            "model_tflops": round(wl["gflop_per_fwd"] * a.batch / ms, 1),
            "rel_l2_vs_fp32_strict": float(O.rel_err(out.cpu(), ref.cpu())),
            "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}), flush=True)


if __name__ == "__main__":
    main()
