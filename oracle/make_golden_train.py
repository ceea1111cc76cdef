"""Generates tests/golden/train_*.npz by running the UNMODIFIED reference's ``GaussianDiffusion.training_losses``
(training_project/utils/gaussian_diffusion.py:824-983) on the small DS-Diff backbone, CPU fp32.  Run here only (needs
/root/reference):  python oracle/make_golden_train.py
The only harness patch is ``get_heatmap`` (a matplotlib visualisation of the logits, not part of the loss)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import  # noqa: E402
import make_golden as MG  # noqa: E402
from make_golden import load_synth, rnd, save  # noqa: E402


def main():
    _ref_import.install()
    import training_project.utils.gaussian_diffusion as GD
    from training_project.utils.script_util import create_gaussian_diffusion
    from UNet_DS_Diff.model import DSUnetModel
    GD.get_heatmap = lambda m: m

    class Wrap(torch.nn.Module):      # DiffusionWrapper 'concat' (ddpm.py:1331-1333)
        def __init__(self, net):
            super().__init__()
            self.diffusion_model = net

        def forward(self, x, t, c_concat=None):
            return self.diffusion_model(torch.cat([x] + c_concat, dim=1), t)

    torch.manual_seed(0)
    with torch.no_grad():
        net = DSUnetModel(**MG.DS_FILM2)
        load_synth(net)
    model = Wrap(net)
    B = 3
    x0 = torch.clamp(rnd(31, B, 1, 32, 32) * 0.5, -1, 1)
    x0[0, 0, :4] = -1.0                  # exercises the x < -0.999 / x > 0.999 branches of the decoder NLL
    x0[0, 0, 4:8] = 1.0
    cond = [torch.clamp(rnd(32 + i, B, 1, 32, 32) * 0.5, -1, 1) for i in range(3)]
    noise = rnd(40, B, 1, 32, 32)
    t = torch.tensor([0, 417, 999])      # t == 0 takes the decoder-NLL branch of the VB term
    for tag, rls in (("mse", False), ("rescaled", True)):
        diff = create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                         rescale_learned_sigmas=rls, timestep_respacing="", parameterization="v")
        with torch.no_grad():
            terms = diff.training_losses(model, x0, t, noise=noise, disentangle="eu",
                                         model_kwargs={"F_Data1": cond[0], "F_Data2": cond[1], "S_Data1": cond[2]})
            x_t = diff.q_sample(x0, t, noise=noise)
        save(f"train_ds_film2_{tag}", x0=x0, cond=torch.stack(cond), noise=noise, t=t, x_t=x_t,
             **{k: v for k, v in terms.items() if k != "contrast_map"})

    # gradients of the trainer's total loss (trainers/trainer_use_gaussian_diff.py:464-478):
    #   mean(contrast_lambda * (disen_c_s + disen_s_a_l) + mse + vb)
    lam = 0.1
    diff = create_gaussian_diffusion(steps=1000, learn_sigma=True, noise_schedule="linear", rescale_timesteps=True,
                                     timestep_respacing="", parameterization="v")
    net.zero_grad()
    terms = diff.training_losses(model, x0, t, noise=noise, disentangle="eu", disen_lambda=lam,
                                 model_kwargs={"F_Data1": cond[0], "F_Data2": cond[1], "S_Data1": cond[2]})
    total = (lam * (terms["disen_c_s_loss"] + terms["disen_s_a_l_loss"]) + terms["loss"]).mean()
    total.backward()
    names = [k for k, p in net.named_parameters()]
    norms = np.array([float(p.grad.double().norm()) if p.grad is not None else -1.0 for _, p in net.named_parameters()])
    picks = ["out.2.weight", "time_embed.0.bias", "input_blocks.0.0.weight", "input_blocks_a.0.0.weight",
             "middle_block.0.in_layers.2.bias"]
    params = dict(net.named_parameters())
    save("train_ds_film2_grads", total=total.detach(), lam=np.float64(lam), grad_norms=norms,
         **{"g:" + k: params[k].grad for k in picks})
    with open(os.path.join(MG.GOLD, "train_ds_film2_grads.names.txt"), "w") as f:
        f.write("\n".join(names) + "\n")


if __name__ == "__main__":
    main()
