"""GPU parity of the latent-path components against goldens generated from the
UNMODIFIED reference (oracle/make_golden_ext.py) and against the oracle with operands rounded at the same places.

Tolerances as in test_gpu_model.py: 16-bit tensor-core operands put the per-forward relative L2 error at the 1e-3
level (fp16) / 1e-2 level (bf16)."""
import math

import pytest
import torch
import torch.nn.functional as F

import dsdiff_oracle as O
from _util import golden, maxabs, synth_sd

pytestmark = pytest.mark.gpu
REL_TOL = {torch.float16: 3e-3, torch.bfloat16: 2.5e-2}


def _pkg():
    import diffusion_models_dsdiff_b200 as P
    return P


# -------------------------------------------------------------------------------------------------
# op level: kernels added for the autoencoder
# -------------------------------------------------------------------------------------------------
def test_softmax_rows_conv1x1_gaussian():
    from diffusion_models_dsdiff_b200 import ops
    g = torch.Generator().manual_seed(5)
    s = torch.randn(3, 7, 260, generator=g) * 20
    for dt in (torch.float16, torch.bfloat16):
        p = ops.softmax_rows(s.cuda(), 0.125, dt).float().cpu()
        ref = torch.softmax(s * 0.125, -1)
        assert (p - ref).abs().max().item() < (2e-3 if dt == torch.float16 else 8e-3)
    x = torch.randn(2, 6, 9, 11, generator=g)
    w, b = torch.randn(5, 6, 1, 1, generator=g), torch.randn(5, generator=g)
    got = ops.conv1x1_f32(x.cuda(), w.cuda(), b.cuda()).cpu()
    assert (got - F.conv2d(x, w, b)).abs().max().item() < 1e-5
    mom = torch.randn(2, 6, 8, 8, generator=g) * 3
    noise = torch.randn(2, 3, 8, 8, generator=g)
    ref = O.ae_posterior_sample(mom, noise)
    assert (ops.gaussian_sample(mom.cuda(), noise.cuda()).cpu() - ref).abs().max().item() < 1e-5
    assert torch.equal(ops.gaussian_sample(mom.cuda(), None).cpu(), mom[:, :3])


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_conv_fp32_nhwc_output_and_asymmetric_pad(dt):
    from diffusion_models_dsdiff_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 64, 16, 16, generator=g).to(dt).float()
    w = (torch.randn(96, 64, 1, 1, generator=g) / 8).to(dt).float()
    xa = x.permute(0, 2, 3, 1).contiguous().to(dt).cuda()
    y = ops.conv2d([(xa, ops.pack_conv_weight(w.cuda(), dt), 0, 1, 0)], out_f32=True)
    assert y.dtype == torch.float32
    ref = F.conv2d(x, w).permute(0, 2, 3, 1)
    assert (y.cpu() - ref).abs().max().item() < 1e-4                      # fp32 straight from the accumulators
    # Downsample of the KL autoencoder: pad (0,1,0,1) then 3x3 stride 2 pad 0 (model.py:80-85)
    w3 = (torch.randn(64, 64, 3, 3, generator=g) / 24).to(dt).float()
    ref = F.conv2d(F.pad(x, (0, 1, 0, 1)), w3, stride=2)
    y = ops.conv2d([(xa, ops.pack_conv_weight(w3.cuda(), dt), 0, 3, 0)], stride=2, out_hw=(8, 8))
    err = (y.float().cpu().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err < (4e-3 if dt == torch.float16 else 3e-2)


# -------------------------------------------------------------------------------------------------
# a17: AutoencoderKL encode / decode
# -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_autoencoder_kl_matches_reference(dtype):
    P = _pkg()
    hp, sd = synth_sd("ae_small")
    ae = P.AutoencoderKL(hp["ddconfig"], None, hp["embed_dim"], act_dtype=dtype)
    assert set(ae.state_dict().keys()) == set(sd.keys())
    ae.load_state_dict(sd, strict=True)
    ae = ae.cuda().eval()
    g = golden("ae_small")
    post = ae.encode(g["x"].cuda())
    assert post.parameters.shape == g["moments"].shape
    assert O.rel_err(post.parameters.cpu(), g["moments"]) < REL_TOL[dtype]
    # against the oracle with operands rounded at the same places
    momq = O.ae_encode_moments(sd, g["x"], opq=lambda v: v.to(dtype).float())
    assert O.rel_err(post.parameters.cpu(), momq) < REL_TOL[dtype]
    z = post.sample(noise=g["noise"].cuda())
    assert maxabs(z.cpu(), O.ae_posterior_sample(post.parameters.cpu(), g["noise"])) < 1e-5
    dec = ae.decode(g["z"].cuda())
    assert dec.shape == g["dec"].shape and dec.dtype == torch.float32
    assert O.rel_err(dec.cpu(), g["dec"]) < REL_TOL[dtype]
    assert maxabs(dec.cpu(), g["dec"]) < (1e-2 if dtype == torch.float16 else 8e-2) * max(1.0, float(g["dec"].abs().max()))


def test_autoencoder_cpu_raises():
    P = _pkg()
    hp, sd = synth_sd("ae_small")
    ae = P.AutoencoderKL(hp["ddconfig"], None, hp["embed_dim"])
    with pytest.raises(P.B200Error):
        ae.decode(torch.zeros(1, 3, 16, 16))
