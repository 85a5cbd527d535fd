"""ctypes binding of libfrido_b200.so (include/frido_b200.h).

There is NO fallback: if the shared library is missing or a launcher returns an
error the caller gets an exception.  The structures below mirror the header
field for field; `frido_sizeof_op()` is checked at load time.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfrido_b200.so")

c_f = C.c_float
c_i = C.c_int32
c_l = C.c_int64
c_p = C.c_void_p

ACT_NONE, ACT_RELU, ACT_SILU, ACT_GEGLU, ACT_GEGLU_FAST, ACT_GELU = 0, 1, 2, 3, 4, 5
(OP_CONV, OP_GN_STATS, OP_NORM_ACT, OP_LAYERNORM, OP_SOFTMAX, OP_TIME_EMBED, OP_STEP_BEGIN, OP_UPDATE, OP_SNAP, OP_VQ,
 OP_ZERO, OP_UPSAMPLE, OP_EMBED, OP_MHA, OP_CONVT, OP_ASSEMBLE, OP_ATTN, OP_BLEND, OP_GN_FINALIZE, OP_FLASH) = range(1, 21)


class ConvParams(C.Structure):
    _fields_ = [
        ("a0", c_p), ("a1", c_p), ("c0", c_i), ("c1", c_i),
        ("a0_sb", c_l), ("a0_sy", c_l), ("a0_sx", c_l), ("a0_sc", c_l),
        ("a1_sb", c_l), ("a1_sy", c_l), ("a1_sx", c_l), ("a1_sc", c_l),
        ("B", c_i), ("Hin", c_i), ("Win", c_i), ("ups", c_i),
        ("ksize", c_i), ("stride", c_i), ("pad", c_i), ("Hout", c_i), ("Wout", c_i),
        ("w", c_p), ("w_sb", c_l), ("w_ld", c_l), ("w_lo", c_p), ("Cout", c_i),
        ("bias", c_p), ("rowvec", c_p), ("rowvec_sb", c_l), ("res", c_p),
        ("alpha", c_f), ("act", c_i), ("out", c_p),
        ("o_sb", c_l), ("o_sp", c_l), ("o_sn", c_l), ("round_tf32", c_i), ("out_hi", c_p), ("out_lo", c_p), ("chan_sums", c_p),
        ("x0", c_p), ("x1", c_p), ("cx0", c_i), ("cx1", c_i), ("x0_sb", c_l), ("x0_sy", c_l), ("x0_sx", c_l),
        ("x1_sb", c_l), ("x1_sy", c_l), ("x1_sx", c_l),
        ("nrm_ab", c_p), ("nrm_gb", c_p), ("nrm_silu", c_i), ("a_presplit", c_i), ("out_u8", c_p), ("u8_mode", c_i), ("sk_ws", c_p), ("sk_ws_bytes", c_l), ("engine", c_i),
    ]


class GnStatsParams(C.Structure):
    _fields_ = [("a0", c_p), ("a1", c_p), ("c0", c_i), ("c1", c_i), ("B", c_i), ("HW", c_i), ("groups", c_i),
                ("sums", c_p)]


class NormActParams(C.Structure):
    _fields_ = [("a0", c_p), ("a1", c_p), ("c0", c_i), ("c1", c_i), ("B", c_i), ("HW", c_i), ("groups", c_i),
                ("sums", c_p), ("eps", c_f), ("csum0", c_p), ("csum1", c_p), ("gamma", c_p), ("beta", c_p), ("gb", c_p), ("silu", c_i),
                ("round_tf32", c_i), ("out", c_p), ("out_split", c_i)]


class LayerNormParams(C.Structure):
    _fields_ = [("x", c_p), ("rows", c_l), ("C", c_i), ("eps", c_f), ("gamma", c_p), ("beta", c_p),
                ("round_tf32", c_i), ("out", c_p), ("out_hi", c_p), ("out_lo", c_p)]


class SoftmaxParams(C.Structure):
    _fields_ = [("s", c_p), ("rows", c_l), ("n", c_i), ("ld", c_l), ("scale", c_f), ("round_tf32", c_i), ("out", c_p)]


class TimeEmbedParams(C.Structure):
    _fields_ = [("t", c_p), ("B", c_i), ("dim", c_i), ("max_period", c_f), ("freqs", c_p), ("out", c_p)]


class StepBeginParams(C.Structure):
    _fields_ = [("step", c_p), ("t_table", c_p), ("use_next", c_i), ("T", c_i), ("ts", c_p), ("B", c_i)]


class UpdateParams(C.Structure):
    _fields_ = [
        ("x", c_p), ("eps", c_p), ("eps_uncond", c_p), ("cfg_scale", c_f),
        ("B", c_i), ("c_start", c_i), ("c_end", c_i), ("HW", c_i),
        ("coef", c_p), ("step", c_p), ("advance", c_i), ("plms_order", c_i), ("plms_mode", c_i),
        ("hist", c_p), ("eps_save", c_p), ("noise", c_p), ("seed", C.c_uint64), ("seed_dev", c_p), ("temperature", c_f),
        ("x_prev", c_p), ("x_dup", c_p), ("pred_x0", c_p),
    ]


class SnapParams(C.Structure):
    _fields_ = [("x", c_p), ("B", c_i), ("C", c_i), ("H", c_i), ("W", c_i), ("c_start", c_i), ("c_end", c_i), ("n", c_i)]


class VqParams(C.Structure):
    _fields_ = [("z", c_p), ("B", c_i), ("C_total", c_i), ("HW", c_i), ("c_start", c_i), ("e_dim", c_i),
                ("scale_factor", c_f), ("codebook", c_p), ("n_e", c_i), ("out", c_p), ("out_C", c_i),
                ("out_coff", c_i), ("indices", c_p), ("z_nhwc", c_i)]


class ZeroParams(C.Structure):
    _fields_ = [("ptr", c_p), ("nbytes", c_l)]


class UpsampleParams(C.Structure):
    _fields_ = [("x", c_p), ("B", c_i), ("H", c_i), ("W", c_i), ("C", c_i), ("round_tf32", c_i), ("out", c_p), ("out_split", c_i)]


class EmbedParams(C.Structure):
    _fields_ = [("tokens", c_p), ("B", c_i), ("L", c_i), ("D", c_i), ("vocab", c_i), ("tok_emb", c_p), ("pos_emb", c_p), ("out", c_p)]


class MhaParams(C.Structure):
    _fields_ = [("qkv", c_p), ("B", c_i), ("L", c_i), ("H", c_i), ("Dh", c_i), ("scale", c_f), ("out", c_p)]


class AttnParams(C.Structure):
    _fields_ = [("q", c_p), ("q_sb", c_l), ("q_ld", c_l), ("k", c_p), ("k_sb", c_l), ("k_ld", c_l), ("v", c_p), ("v_sb", c_l), ("v_ld", c_l),
                ("B", c_i), ("N", c_i), ("Nk", c_i), ("C", c_i), ("scale", c_f), ("out", c_p), ("o_sb", c_l), ("o_ld", c_l),
                ("ln_gamma", c_p), ("ln_beta", c_p), ("ln_eps", c_f), ("bias", c_p), ("res", c_p), ("r_sb", c_l), ("r_ld", c_l),
                ("ln2_gamma", c_p), ("ln2_beta", c_p), ("ln2_eps", c_f), ("out2", c_p)]


class ConvT2dParams(C.Structure):
    _fields_ = [("x", c_p), ("B", c_i), ("H", c_i), ("W", c_i), ("Cin", c_i), ("Cout", c_i), ("w", c_p), ("bias", c_p), ("out", c_p), ("x_ld", c_i)]


class AssembleParams(C.Structure):
    _fields_ = [("h", c_p), ("B", c_i), ("H", c_i), ("W", c_i), ("e", c_i), ("sh", c_i), ("scale", c_f), ("out", c_p),
                ("C_total", c_i), ("c_off", c_i)]


class ToU8Params(C.Structure):
    _fields_ = [("x", c_p), ("B", c_i), ("C", c_i), ("HW", c_i), ("mode", c_i), ("out", c_p)]


class BlendParams(C.Structure):
    _fields_ = [("x", c_p), ("x_dup", c_p), ("x0", c_p), ("mask", c_p), ("noise", c_p), ("sqrt_acp", c_p), ("sqrt_1m_acp", c_p),
                ("step", c_p), ("t_table", c_p), ("T", c_i), ("B", c_i), ("C", c_i), ("HW", c_i), ("seed", C.c_uint64),
                ("seed_dev", c_p)]


class GnFinalizeParams(C.Structure):
    _fields_ = [("c0", c_i), ("c1", c_i), ("B", c_i), ("HW", c_i), ("groups", c_i), ("eps", c_f), ("sums", c_p), ("csum0", c_p),
                ("csum1", c_p), ("gamma", c_p), ("beta", c_p), ("ab", c_p)]


class FlashParams(C.Structure):
    _fields_ = [("q_hi", c_p), ("q_lo", c_p), ("q_sb", c_l), ("q_ld", c_l), ("k_hi", c_p), ("k_lo", c_p), ("k_sb", c_l), ("k_ld", c_l),
                ("vt_hi", c_p), ("vt_lo", c_p), ("vt_sb", c_l), ("vt_ld", c_l), ("B", c_i), ("N", c_i), ("C", c_i), ("scale", c_f),
                ("bias", c_p), ("res", c_p), ("r_sb", c_l), ("r_ld", c_l), ("out", c_p), ("o_sb", c_l), ("o_ld", c_l)]


class _OpU(C.Union):
    _fields_ = [("conv", ConvParams), ("gn_stats", GnStatsParams), ("norm_act", NormActParams),
                ("layernorm", LayerNormParams), ("softmax", SoftmaxParams), ("time_embed", TimeEmbedParams),
                ("step_begin", StepBeginParams), ("update", UpdateParams), ("snap", SnapParams), ("vq", VqParams),
                ("zero", ZeroParams), ("upsample", UpsampleParams), ("embed", EmbedParams), ("mha", MhaParams), ("convt", ConvT2dParams),
                ("assemble", AssembleParams), ("attn", AttnParams), ("blend", BlendParams), ("gn_finalize", GnFinalizeParams), ("flash", FlashParams)]


class Op(C.Structure):
    _fields_ = [("kind", c_i), ("tag", c_i), ("u", _OpU)]


_KIND_FIELD = {OP_CONV: "conv", OP_GN_STATS: "gn_stats", OP_NORM_ACT: "norm_act", OP_LAYERNORM: "layernorm",
               OP_SOFTMAX: "softmax", OP_TIME_EMBED: "time_embed", OP_STEP_BEGIN: "step_begin", OP_UPDATE: "update",
               OP_SNAP: "snap", OP_VQ: "vq", OP_ZERO: "zero", OP_UPSAMPLE: "upsample", OP_EMBED: "embed", OP_MHA: "mha", OP_CONVT: "convt", OP_ASSEMBLE: "assemble", OP_ATTN: "attn",
               OP_BLEND: "blend", OP_GN_FINALIZE: "gn_finalize", OP_FLASH: "flash"}

EXPORTS = [
    "frido_conv2d", "frido_gn_stats", "frido_norm_act", "frido_layernorm", "frido_softmax", "frido_time_embed",
    "frido_step_begin", "frido_sampler_update", "frido_stage_snap", "frido_vq_lookup", "frido_zero",
    "frido_round_tf32", "frido_split_bf16", "frido_upsample2x", "frido_to_uint8", "frido_embed_tokens", "frido_mha_small", "frido_attn_small", "frido_conv_transpose2d", "frido_assemble_latent", "frido_mask_blend", "frido_gn_finalize", "frido_attn_flash", "frido_attn_flash_eligible", "frido_pack_permute3", "frido_pack_conv_weight", "frido_matmul_f64acc",
    "frido_fold_self_attention", "frido_vec_add", "frido_workspace_bytes", "frido_run_program", "frido_abi_version", "frido_sizeof_op", "frido_last_error",
    "frido_launch_count", "frido_check_device",
]

SK_WS_BYTES = 40 << 20
ABI_VERSION = 7

_lib = None


class FridoError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FridoError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C frido_b200/csrc). There is no CPU or PyTorch fallback."
        )
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise FridoError(f"{LIB_PATH} does not export {name}")
    L.frido_last_error.restype = C.c_char_p
    L.frido_launch_count.restype = C.c_int64
    L.frido_run_program.argtypes = [C.c_void_p, c_i, c_p]
    L.frido_zero.argtypes = [c_p, c_l, c_p]
    L.frido_round_tf32.argtypes = [c_p, c_p, c_l, c_p]
    L.frido_split_bf16.argtypes = [c_p, c_p, c_p, c_l, c_p]
    L.frido_attn_flash_eligible.argtypes = [c_i, c_i, c_i]
    L.frido_pack_permute3.argtypes = [c_p, c_l, c_l, c_l, c_p, c_l, c_l, c_l, c_i, c_i, c_i, c_p]
    L.frido_pack_conv_weight.argtypes = [c_p, c_i, c_i, c_i, c_i, c_p, c_l, c_p]
    L.frido_matmul_f64acc.argtypes = [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_i, c_i, c_p, c_l, c_p]
    L.frido_fold_self_attention.argtypes = [c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p]
    L.frido_vec_add.argtypes = [c_p, c_p, c_p, c_l, c_p]
    L.frido_workspace_bytes.argtypes = [c_p, c_i]
    L.frido_workspace_bytes.restype = C.c_int64
    for name in ("frido_conv2d", "frido_gn_stats", "frido_norm_act", "frido_layernorm", "frido_softmax",
                 "frido_time_embed", "frido_step_begin", "frido_sampler_update", "frido_stage_snap", "frido_vq_lookup",
                 "frido_upsample2x", "frido_to_uint8", "frido_embed_tokens", "frido_mha_small", "frido_attn_small",
                 "frido_conv_transpose2d", "frido_assemble_latent", "frido_mask_blend", "frido_gn_finalize", "frido_attn_flash"):
        getattr(L, name).argtypes = [c_p, c_p]
    if L.frido_abi_version() != ABI_VERSION:
        raise FridoError(f"ABI mismatch: library version {L.frido_abi_version()}, binding {ABI_VERSION} (rebuild: make -C frido_b200/csrc)")
    if L.frido_sizeof_op() != C.sizeof(Op):
        raise FridoError(f"ABI mismatch: sizeof(FridoOp) C={L.frido_sizeof_op()} python={C.sizeof(Op)}")
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().frido_last_error().decode(errors="replace")
        raise FridoError(f"{what} failed (rc={rc}): {msg}")


def make_op(kind, params, tag=0):
    op = Op()
    op.kind = kind
    op.tag = tag
    setattr(op.u, _KIND_FIELD[kind], params)
    return op
