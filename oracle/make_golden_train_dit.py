"""Generates tests/golden/train_dit_*.npz: ONE TRAINING STEP of the reference — its own ``DiT`` (UNet_DS_Diff/DiT_models.py, on
the timm restatement of make_golden_ext.py) behind a concat wrapper, its own ``GaussianDiffusion.training_losses``
(training_project/utils/gaussian_diffusion.py:824-983), ``loss.backward()`` and ``torch.optim.AdamW(params, lr)``
(trainers/trainer_use_gaussian_diff_adv.py:306-320, 345-382) — on CPU fp32.  Run here only (needs /root/reference):

    python oracle/make_golden_train_dit.py

Fixtures: the loss terms, the gradient of every parameter (norms for all, full tensors for a few) and the parameters after
two optimizer steps.  They are the parity anchor of diffusion_models_dsdiff_b200/dit_train.py."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402
import make_golden as MG  # noqa: E402
import make_golden_ext as MGE  # noqa: E402
from make_golden import load_synth, rnd, save  # noqa: E402

DIT_TRAIN = dict(input_size=32, patch_size=4, in_channels=4, hidden_size=128, depth=2, num_heads=4, learn_sigma=True)
DIT_TRAIN64 = dict(input_size=16, patch_size=2, in_channels=4, hidden_size=128, depth=2, num_heads=2, learn_sigma=True)
PICKS = ["final_layer.linear.weight", "final_layer.linear.bias", "blocks.1.mlp.fc2.weight", "blocks.1.mlp.fc1.bias",
         "blocks.0.attn.qkv.weight", "blocks.0.attn.qkv.bias", "blocks.0.attn.proj.weight",
         "blocks.0.adaLN_modulation.1.weight", "blocks.1.adaLN_modulation.1.bias", "final_layer.adaLN_modulation.1.weight",
         "x_embedder.proj.weight", "x_embedder.proj.bias", "t_embedder.mlp.0.weight", "t_embedder.mlp.2.bias"]


def main():
    _ref_import.install()
    MGE._install_timm_restatement()
    import training_project.utils.gaussian_diffusion as GD
    from training_project.utils.script_util import create_gaussian_diffusion
    from UNet_DS_Diff.DiT_models import DiT
    GD.get_heatmap = lambda m: m

    class Wrap(torch.nn.Module):      # DiffusionWrapper 'concat' (ddpm.py:1331-1333)
        def __init__(self, net):
            super().__init__()
            self.diffusion_model = net

        def forward(self, x, t, c_concat=None):
            return self.diffusion_model(torch.cat([x] + c_concat, dim=1), t)

    entries = {}
    for name, hp, param, rls in (("train_dit_v", DIT_TRAIN, "v", False), ("train_dit_eps_rescaled", DIT_TRAIN64, "eps", True)):
        torch.manual_seed(0)
        net = DiT(**hp)
        pos_ref = net.pos_embed.detach().clone()
        shapes = load_synth(net)
        net.pos_embed.data.copy_(pos_ref)
        net.train()                      # no dropout anywhere in this configuration; y is None
        model = Wrap(net)
        s = hp["input_size"]
        B = 3
        x0 = torch.clamp(rnd(61, B, 1, s, s) * 0.5, -1, 1)
        x0[0, 0, :2] = -1.0               # exercises the x < -0.999 / x > 0.999 branches of the decoder NLL (t == 0)
        x0[0, 0, 2:4] = 1.0
        cond = [torch.clamp(rnd(62 + i, B, 1, s, s) * 0.5, -1, 1) for i in range(3)]
        noise = rnd(70, B, 1, s, s)
        t = torch.tensor([0, 417, 999])
        diff = create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                         rescale_learned_sigmas=rls, timestep_respacing="", parameterization=param)
        kw = {"F_Data1": cond[0], "F_Data2": cond[1], "S_Data1": cond[2]}
        lr = 1e-3
        opt = torch.optim.AdamW([p for p in net.parameters()], lr=lr)
        net.zero_grad()
        terms = diff.training_losses(model, x0, t, noise=noise, model_kwargs=kw)
        total = terms["loss"].mean()
        total.backward()
        names = [k for k, p in net.named_parameters() if p.requires_grad]
        params = dict(net.named_parameters())
        norms = np.array([float(params[k].grad.double().norm()) if params[k].grad is not None else -1.0 for k in names])
        arrs = dict(x0=x0, cond=torch.stack(cond), noise=noise, t=t, loss=terms["loss"].detach(), mse=terms["mse"].detach(),
                    vb=terms["vb"].detach(), total=total.detach(), grad_norms=norms, lr=np.float64(lr), pos_embed=pos_ref)
        arrs.update({"g:" + k: params[k].grad for k in PICKS})
        # two optimizer steps (the second one on fresh noise / timesteps) -> parameters afterwards
        opt.step()
        noise2, t2 = rnd(71, B, 1, s, s), torch.tensor([5, 250, 761])
        net.zero_grad()
        terms2 = diff.training_losses(model, x0, t2, noise=noise2, model_kwargs=kw)
        terms2["loss"].mean().backward()
        opt.step()
        arrs.update(noise2=noise2, t2=t2, loss2=terms2["loss"].detach())
        arrs.update({"p2:" + k: params[k].detach() for k in PICKS})
        save(name, **arrs)
        with open(os.path.join(MG.GOLD, name + ".names.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
        entries[name] = {"hp": hp, "shapes": {k: v for k, v in shapes.items() if k != "pos_embed"},
                         "parameterization": param, "rescaled": rls}
    MGE.update_manifest(entries)


if __name__ == "__main__":
    main()
