"""ctypes binding of libb200diff.so (include/b200diff.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, the product
raises.  Build it with ``python -c 'import __graft_entry__ as g; g.build()'`` (or ``make`` in
``diffusion_models_dsdiff_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200DIFF_LIB: another build of the same ABI (kernel A/B timing); default = the in-tree library next to this file
LIB_PATH = os.environ.get("B200DIFF_LIB") or os.path.join(_HERE, "libb200diff.so")

ABI_VERSION = 16        # must equal B200_ABI_VERSION of the header the library was built from
F16, BF16 = 0, 1
SAMPLER_LDM_DDIM, SAMPLER_LDM_DDPM, SAMPLER_GD_DDIM, SAMPLER_GD_DDPM, SAMPLER_DPMPP_2M = 0, 1, 2, 3, 4
SAMPLER_NCOEF = 16
PARAM_EPS, PARAM_X0, PARAM_V = 0, 1, 2


class B200Error(RuntimeError):
    pass


class ConvSeg(C.Structure):
    _fields_ = [("src", C.c_void_p), ("weight", C.c_void_p), ("cin", C.c_int), ("wcin", C.c_int),
                ("coff", C.c_int), ("ksize", C.c_int), ("pad", C.c_int), ("pad_w1", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("N", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Hout", C.c_int), ("Wout", C.c_int),
                ("Cout", C.c_int), ("stride", C.c_int), ("nseg", C.c_int), ("seg", ConvSeg * 3),
                ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("rowbias_stride", C.c_int),
                ("residual", C.c_void_p), ("res_scale", C.c_float), ("out", C.c_void_p),
                ("gn_sums", C.c_void_p), ("act_dtype", C.c_int), ("out_nchw_f32", C.c_void_p), ("cout_real", C.c_int), ("rowscale", C.c_void_p), ("rowscale_stride", C.c_int), ("epi_act", C.c_int), ("out_nhwc_f32", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("range_events", C.c_void_p),
                ("wsets", C.c_int), ("rowbias_gstride", C.c_int), ("rowscale_gstride", C.c_int),
                ("out_sy", C.c_int), ("out_sx", C.c_int), ("out_oy", C.c_int), ("out_ox", C.c_int),
                ("force_ksplit", C.c_int), ("out_lo", C.c_void_p), ("residual_lo", C.c_void_p)]


class SamplerDesc(C.Structure):
    _fields_ = [("mode", C.c_int), ("param", C.c_int), ("clip_denoised", C.c_int), ("learned_range", C.c_int),
                ("N", C.c_int), ("C", C.c_int), ("HW", C.c_int), ("x", C.c_void_p), ("model_out", C.c_void_p),
                ("noise", C.c_void_p), ("coef", C.c_void_p), ("step_ptr", C.c_void_p), ("step_host", C.c_int),
                ("x_prev", C.c_void_p), ("pred_x0", C.c_void_p), ("hist", C.c_void_p), ("thresh", C.c_void_p),
                ("phase", C.c_int)]


class AttnDesc(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
                ("B", C.c_int), ("Lq", C.c_int), ("Lk", C.c_int), ("heads", C.c_int), ("d", C.c_int),
                ("ldq", C.c_int), ("ldk", C.c_int), ("ldv", C.c_int), ("ldo", C.c_int),
                ("q_off", C.c_int), ("k_off", C.c_int), ("v_off", C.c_int),
                ("q_hstride", C.c_int), ("k_hstride", C.c_int), ("v_hstride", C.c_int),
                ("scale", C.c_float), ("act_dtype", C.c_int), ("force_simt", C.c_int), ("lse", C.c_void_p)]


_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64
# name -> argtypes (all return int unless listed in _RESTYPE); mirrors include/b200diff.h
SIGNATURES = {
    "b200_abi_version": [],
    "b200_last_error": [],
    "b200_init": [],
    "b200_launch_count": [],
    "b200_nchw_f32_to_nhwc": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_nhwc_to_nchw_f32": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_pack_conv_weight": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_conv2d_igemm": [C.POINTER(ConvDesc), _P],
    "b200_conv_ksplit": [C.POINTER(ConvDesc)],
    "b200_conv3x3_in": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P],
    "b200_conv3x3_out": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_groupnorm_stats": [_P, _I, _P, _I, _I, _I, _I, _P, _I, _P],
    "b200_groupnorm_apply": [_P, _I, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P, _I, _I, _P, _I, _P],
    "b200_channel_stats": [_P, _I, _I, _I, _P, _I, _P],
    "b200_groupnorm_apply_cs": [_P, _I, _P, _P, _I, _P, _I, _I, _I, _F, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P],
    "b200_attention_qkv": [_P, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    "b200_attention": [C.POINTER(AttnDesc), _P],
    "b200_layernorm": [_P, _P, _P, _P, _P, _I, _I, _L, _I, _F, _P, _I, _P],
    "b200_geglu": [_P, _P, _L, _I, _I, _P],
    "b200_patchify": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "b200_unpatchify": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "b200_add_f32": [_P, _P, _P, _L, _P],
    "b200_cfg_combine": [_P, _P, _F, _P, _L, _P],
    "b200_softmax_rows": [_P, _P, _L, _I, _F, _I, _P],
    "b200_conv1x1_f32_nchw": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "b200_gaussian_sample": [_P, _P, _P, _I, _I, _I, _P],
    "b200_timestep_embedding": [_P, _P, _P, _P, _I, _I, _P],
    "b200_linear_f32": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "b200_silu_f32": [_P, _P, _L, _P],
    "b200_gather_row_f32": [_P, _P, _P, _I, _P],
    "b200_upsample_nearest2x": [_P, _P, _I, _I, _I, _I, _I, _P],
    "b200_avgpool2x": [_P, _P, _I, _I, _I, _I, _I, _P],
    "b200_head_gather9": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "b200_add4_scale": [_P, _P, _P, _P, _F, _P, _L, _I, _I, _P],
    "b200_add4_scale_stats": [_P, _P, _P, _P, _F, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P],
    "b200_se_attention": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "b200_sampler_update": [C.POINTER(SamplerDesc), _P],
    "b200_abs_quantile": [_P, _I, _L, _F, _F, _P, _P],
    "b200_lincomb2": [_P, _P, _P, _P, _P, _I, _L, _I, _P],
    "b200_advance_step": [_P, _I, _P],
    # training step (csrc/train.cu)
    "b200_transpose16": [_P, _P, _P, _L, _I, _I, _P],
    "b200_colsum16": [_P, _P, _L, _I, _I, _P],
    "b200_wgrad_splits": [_L, _I, _I],
    "b200_linear_wgrad": [_P, _P, _P, _P, _L, _I, _I, _I, _P],
    "b200_gelu_fwd": [_P, _P, _L, _I, _P],
    "b200_gelu_bwd": [_P, _P, _P, _L, _I, _P],
    "b200_gated_residual": [_P, _P, _P, _I, _P, _L, _I, _I, _I, _P],
    "b200_gate_bwd": [_P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P],
    "b200_ln_mod_bwd": [_P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    "b200_attention_bwd": [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _P],
    "b200_unpatchify_bwd": [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "b200_q_sample": [_P, _P, _P, _P, _I, _I, _P],
    "b200_diffusion_loss": [_P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _P],
    "b200_linear_f32_wgrad": [_P, _I, _P, _I, _I, _P, _P, _I, _I, _I, _I, _P],
    "b200_linear_f32_dgrad": [_P, _I, _P, _P, _P, _I, _I, _I, _P],
    "b200_adamw": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P, _P],
}
_RESTYPE = {"b200_last_error": C.c_char_p, "b200_launch_count": C.c_uint64}

_lib = None
_inited_devices = set()


def lib() -> C.CDLL:
    """The loaded library; raises (no fallback) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run __graft_entry__.build()).  There is no CPU / eager fallback.")
        l = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, C.c_int)
        got = int(l.b200_abi_version())
        if got != ABI_VERSION:      # struct layouts below would not match: refuse instead of corrupting memory
            raise B200Error(f"{LIB_PATH} was built for ABI {got}, this package binds ABI {ABI_VERSION}: rebuild it "
                            "(__graft_entry__.build() or `make -C diffusion_models_dsdiff_b200/csrc`)")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise B200Error(f"{what or 'libb200diff'} failed (rc={rc}): {lib().b200_last_error().decode()}")


def ensure_init(device_index: int) -> None:
    """b200_init() once per device (must be the current device)."""
    if device_index not in _inited_devices:
        check(lib().b200_init(), "b200_init")
        _inited_devices.add(device_index)


def launch_count() -> int:
    return int(lib().b200_launch_count())
