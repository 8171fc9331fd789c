"""ctypes binding of libtvt_b200.so — one Structure per args struct of include/tvt.h, same field order.

The library is required: there is no CPU or PyTorch fallback.  ``load()`` raises if the shared object has
not been built (``python "data-efficient-video-transformers_b200/build.py"``), and every wrapper raises
``TvtError`` carrying ``tvt_last_error()`` when an entry point returns a negative status.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtvt_b200.so")

TVT_BF16, TVT_F32, TVT_F64 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
MAX_POOL_SCALES = 8
MAX_EXPERTS = 8

vp, i64, i32, f32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_uint64


class TvtError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [("a", vp), ("a_lo", vp), ("b", vp), ("b_lo", vp), ("m", i64), ("n", i64), ("k", i64),
                ("lda", i64), ("ldb", i64), ("a_mn_major", i32), ("b_mn_major", i32), ("splits", i32), ("act", i32),
                ("alpha", f32), ("bias", vp), ("residual", vp), ("residual_dtype", i32), ("ld_residual", i64),
                ("relu_mask", vp), ("mask_dtype", i32), ("ld_mask", i64), ("gelu_gate", vp), ("gate_dtype", i32),
                ("ld_gate", i64), ("dropout_p", f32), ("dropout_seed", u64), ("out_preact", vp),
                ("preact_dtype", i32), ("ld_preact", i64), ("out_f32", vp), ("ld_f32", i64), ("atomic_out", i32),
                ("out_bf16", vp), ("out_bf16_lo", vp), ("ld_bf16", i64),
                ("ln_in_stats", vp), ("ln_in_c", vp), ("ln_res_stats", vp), ("ln_res_gamma", vp), ("ln_res_beta", vp),
                ("stats_out", vp), ("ln_dim", i64), ("ln_eps", f32), ("reserved", i32), ("a_rowsum", vp)]


class LayerNormFwdArgs(C.Structure):
    _fields_ = [("x", vp), ("cls", vp), ("pe", vp), ("gamma", vp), ("beta", vp), ("y", vp), ("pre", vp),
                ("mean", vp), ("rstd", vp), ("rows", i64), ("d", i64), ("seq_len", i64), ("dtype", i32),
                ("eps", f32), ("dropout_p", f32), ("dropout_seed", u64), ("y_seq", i64), ("y_pitch", i64)]


class LayerNormBwdArgs(C.Structure):
    _fields_ = [("dy", vp), ("x", vp), ("mean", vp), ("rstd", vp), ("gamma", vp), ("dx", vp), ("dz", vp),
                ("dfeat", vp), ("dcls", vp), ("dgamma", vp), ("dbeta", vp), ("dbias", vp), ("rows", i64),
                ("d", i64), ("seq_len", i64), ("dtype", i32), ("dropout_p", f32), ("dropout_seed", u64), ("dres", vp)]


class AttentionFwdArgs(C.Structure):
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("o", vp), ("lse", vp), ("batch", i64), ("heads", i64),
                ("sq", i64), ("sk", i64), ("head_dim", i64), ("ldq", i64), ("ldk", i64), ("ldv", i64), ("ldo", i64),
                ("scale", f32), ("dtype", i32), ("impl", i32), ("dropout_p", f32), ("dropout_seed", u64)]


class AttentionBwdArgs(C.Structure):
    _fields_ = [("q", vp), ("k", vp), ("v", vp), ("o", vp), ("d_o", vp), ("lse", vp), ("dq", vp), ("dk", vp),
                ("dv", vp), ("batch", i64), ("heads", i64), ("sq", i64), ("sk", i64), ("head_dim", i64),
                ("ldq", i64), ("ldk", i64), ("ldv", i64), ("ldo", i64), ("lddo", i64), ("lddq", i64), ("lddk", i64),
                ("lddv", i64), ("scale", f32), ("dtype", i32), ("impl", i32), ("dropout_p", f32), ("dropout_seed", u64)]


class PyramidPoolFwdArgs(C.Structure):
    _fields_ = [("x", vp), ("batch", i64), ("frames", i64), ("d", i64), ("x_batch_stride", i64),
                ("x_frame_stride", i64), ("num_scales", i32), ("groups", i32 * MAX_POOL_SCALES),
                ("out", vp * MAX_POOL_SCALES), ("dtype", i32), ("relu", i32)]


class PyramidPoolBwdArgs(C.Structure):
    _fields_ = [("dout", vp * MAX_POOL_SCALES), ("out", vp * MAX_POOL_SCALES), ("dx", vp), ("batch", i64),
                ("frames", i64), ("d", i64), ("dx_batch_stride", i64), ("dx_frame_stride", i64), ("num_scales", i32),
                ("groups", i32 * MAX_POOL_SCALES), ("dtype", i32), ("relu", i32), ("accumulate", i32)]


class SpatialPoolArgs(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("frames", i64), ("channels", i64), ("hw", i64), ("ld_out", i64),
                ("col_offset", i64), ("dtype", i32), ("out_dtype", i32)]


class SpatialPoolBwdArgs(C.Structure):
    _fields_ = [("dpooled", vp), ("dx", vp), ("frames", i64), ("channels", i64), ("hw", i64), ("ld", i64),
                ("col_offset", i64), ("dtype", i32), ("reserved", i32)]


class StretchCastArgs(C.Structure):
    _fields_ = [("x", vp), ("y", vp), ("rows", i64), ("d_in", i64), ("d_out", i64), ("ld_out", i64), ("out_dtype", i32), ("reserved", i32)]


class CollabArgs(C.Structure):
    _fields_ = [("c", vp), ("pc", vp), ("a", vp), ("out", vp), ("dout", vp), ("dc", vp), ("dpc", vp), ("da", vp), ("rows", i64), ("d", i64),
                ("experts", i32), ("dtype", i32)]


class L2NormArgs(C.Structure):
    _fields_ = [("x", vp), ("y", vp), ("inv_norm", vp), ("dy", vp), ("dx", vp), ("rows", i64), ("d", i64), ("eps", f32), ("dtype", i32)]


class DistillLossArgs(C.Structure):
    _fields_ = [("student", vp), ("teacher", vp), ("target", vp), ("losses", vp), ("dlogits", vp), ("batch", i64),
                ("classes", i64), ("w_bce", f32), ("w_ce", f32), ("w_kl", f32), ("temperature", f32), ("grad_scale", f32)]


class PyramidHeadArgs(C.Structure):
    _fields_ = [("z", vp), ("target", vp), ("prob", vp), ("loss", vp), ("dz", vp), ("scales", i64), ("batch", i64),
                ("classes", i64), ("grad_scale", f32)]


class ColsumArgs(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("rows", i64), ("cols", i64), ("ld", i64), ("dtype", i32)]


class SplitArgs(C.Structure):
    _fields_ = [("x", vp), ("hi", vp), ("lo", vp), ("n", i64)]


class Split3Args(C.Structure):
    _fields_ = [("x", vp), ("hi4", vp), ("lo4", vp), ("rows", i64), ("cols", i64), ("ld", i64), ("operand", i32), ("reserved", i32)]


class BiasActArgs(C.Structure):
    _fields_ = [("x", vp), ("bias", vp), ("y", vp), ("rows", i64), ("cols", i64), ("out_dtype", i32), ("act", i32),
                ("dropout_p", f32), ("dropout_seed", u64), ("residual", vp), ("preact", vp)]


class PosencArgs(C.Structure):
    _fields_ = [("x", vp), ("pe", vp), ("y", vp), ("rows", i64), ("d", i64), ("seq_len", i64), ("dtype", i32),
                ("dropout_p", f32), ("dropout_seed", u64)]


class ActBwdArgs(C.Structure):
    _fields_ = [("dy", vp), ("y_or_z", vp), ("dx", vp), ("rows", i64), ("cols", i64), ("dtype", i32), ("act", i32),
                ("dropout_p", f32), ("dropout_seed", u64)]


class OptimStepArgs(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("m", vp), ("v", vp), ("p_hi", vp), ("p_lo", vp), ("n", i64), ("kind", i32), ("step", i32),
                ("lr", f32), ("beta1", f32), ("beta2", f32), ("eps", f32), ("weight_decay", f32), ("momentum", f32), ("grad_scale", f32),
                ("step_dev", vp)]


class HeadLinearFwdArgs(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("b", vp), ("y", vp), ("m", i64), ("k", i64), ("classes", i64), ("dtype", i32)]


class HeadLinearBwdArgs(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("dy", vp), ("dx", vp), ("dw", vp), ("db", vp), ("m", i64), ("k", i64),
                ("classes", i64), ("dtype", i32)]


class ClsSumArgs(C.Structure):
    _fields_ = [("tokens", vp * MAX_EXPERTS), ("out", vp), ("batch", i64), ("seq_len", i64), ("d", i64),
                ("num_experts", i32), ("dtype", i32)]


class FeatureAugmentArgs(C.Structure):
    _fields_ = [("x", vp), ("y", vp), ("rows", i64), ("d_in", i64), ("d_out", i64), ("p_drop", C.c_float), ("p_noise", C.c_float),
                ("noise_std", C.c_float), ("seed", C.c_uint64), ("out_dtype", i32)]


class EvalReadoutArgs(C.Structure):
    _fields_ = [("logits", vp), ("target", vp), ("probs", vp), ("labels", vp), ("pred_bits", vp), ("top1", vp),
                ("batch", i64), ("classes", i64), ("row_offset", i64), ("capacity", i64),
                ("thresholds", C.c_float * 16), ("num_thresholds", i32), ("target_dtype", i32)]


# entry point -> args Structure (every symbol include/tvt.h declares that takes (args*, stream))
ENTRY_POINTS = {
    "tvt_gemm": GemmArgs,
    "tvt_layernorm_fwd": LayerNormFwdArgs,
    "tvt_layernorm_bwd": LayerNormBwdArgs,
    "tvt_attention_fwd": AttentionFwdArgs,
    "tvt_attention_bwd": AttentionBwdArgs,
    "tvt_pyramid_pool_fwd": PyramidPoolFwdArgs,
    "tvt_pyramid_pool_bwd": PyramidPoolBwdArgs,
    "tvt_spatial_pool_fwd": SpatialPoolArgs,
    "tvt_spatial_pool_bwd": SpatialPoolBwdArgs,
    "tvt_distill_loss": DistillLossArgs,
    "tvt_stretch_cast": StretchCastArgs,
    "tvt_collab_mix_fwd": CollabArgs,
    "tvt_collab_mix_bwd": CollabArgs,
    "tvt_collab_gate_fwd": CollabArgs,
    "tvt_collab_gate_bwd": CollabArgs,
    "tvt_l2norm_fwd": L2NormArgs,
    "tvt_l2norm_bwd": L2NormArgs,
    "tvt_pyramid_head": PyramidHeadArgs,
    "tvt_colsum": ColsumArgs,
    "tvt_split_f32": SplitArgs,
    "tvt_split_f32x3": Split3Args,
    "tvt_bias_act_fwd": BiasActArgs,
    "tvt_posenc_fwd": PosencArgs,
    "tvt_act_bwd": ActBwdArgs,
    "tvt_optim_step": OptimStepArgs,
    "tvt_head_linear_fwd": HeadLinearFwdArgs,
    "tvt_head_linear_bwd": HeadLinearBwdArgs,
    "tvt_cls_sum_fwd": ClsSumArgs,
    "tvt_eval_readout": EvalReadoutArgs,
    "tvt_feature_augment": FeatureAugmentArgs,
}
PLAIN_SYMBOLS = ("tvt_last_error", "tvt_version", "tvt_device_check", "tvt_set_seed_source", "tvt_step_counter_advance",
                 "tvt_gemm_ln_fold_supported", "tvt_gemm_rowsum_supported")

_lib = None
launches = 0  # number of kernel-launching entry-point calls made through this module (bench.py reads it)


def load():
    """Load the shared library (once).  Raises TvtError when it is missing: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise TvtError(f"{LIB_PATH} not found: build it with `python \"{os.path.join(_HERE, 'build.py')}\"` "
                       "(the CUDA extension is mandatory; there is no CPU/PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.tvt_last_error.restype = C.c_char_p
    lib.tvt_last_error.argtypes = []
    lib.tvt_version.restype = C.c_int
    lib.tvt_device_check.restype = C.c_int
    lib.tvt_set_seed_source.restype = C.c_int
    lib.tvt_set_seed_source.argtypes = [vp]
    lib.tvt_step_counter_advance.restype = C.c_int
    lib.tvt_step_counter_advance.argtypes = [vp, C.c_int, vp]
    lib.tvt_gemm_ln_fold_supported.restype = C.c_int
    lib.tvt_gemm_ln_fold_supported.argtypes = [i64, i64, i64]
    lib.tvt_gemm_rowsum_supported.restype = C.c_int
    lib.tvt_gemm_rowsum_supported.argtypes = [i64, i64, i64, i32]
    for name, st in ENTRY_POINTS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(st), vp]
    _lib = lib
    return lib


def last_error():
    return load().tvt_last_error().decode(errors="replace")


_profile = None  # list of (name, start event, end event, flops) while profile_step() is active


def _flops(name, a):
    if name == "tvt_gemm":
        return 2.0 * a.m * a.n * a.k * (3 if a.a_lo else 1)
    if name == "tvt_attention_fwd":
        return 4.0 * a.batch * a.heads * a.sq * a.sk * a.head_dim
    if name == "tvt_attention_bwd":
        return 10.0 * a.batch * a.heads * a.sq * a.sk * a.head_dim
    return 0.0


def _detail(name, a):
    """Shape key of a launch for the per-shape table of profile_step()."""
    if name == "tvt_gemm":
        st = "".join(t for t, on in (("b", a.bias), ("R", a.act == 1), ("G", a.act == 2), ("m", a.relu_mask), ("g", a.gelu_gate),
                                      ("d", a.dropout_p > 0), ("r", a.residual), ("p", a.out_preact), ("F", a.out_f32), ("L", a.out_bf16_lo)) if on)
        return (f"{a.m}x{a.n}x{a.k} A:{'MN' if a.a_mn_major else 'K'} B:{'MN' if a.b_mn_major else 'K'} "
                f"planes={2 if a.a_lo else 1} splits={a.splits} [{st}]")
    if name in ("tvt_attention_fwd", "tvt_attention_bwd"):
        return f"B={a.batch} H={a.heads} Sq={a.sq} Sk={a.sk} p={a.dropout_p:.2f}"
    return ""


last_profile_detail = {}  # {(entry point, shape key): {ms, flops, calls}} of the latest profile_step()


def call(name, args, stream):
    """Invoke entry point ``name`` with a filled args Structure on cudaStream_t ``stream`` (int)."""
    global launches
    fn = getattr(load(), name)
    if _profile is None:
        rc = fn(C.byref(args), vp(stream))
    else:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(C.byref(args), vp(stream))
        e1.record()
        _profile.append((name, e0, e1, _flops(name, args), _detail(name, args)))
    if rc != 0:
        raise TvtError(f"{name} failed with status {rc}: {last_error()}")
    launches += 1


def set_seed_source(device_ptr):
    """Register (or clear, with None / 0) the device-resident step counter folded into every dropout seed."""
    rc = load().tvt_set_seed_source(vp(device_ptr or 0))
    if rc != 0:
        raise TvtError(f"tvt_set_seed_source failed with status {rc}: {last_error()}")


def step_counter_advance(device_ptr, count, stream):
    global launches
    rc = load().tvt_step_counter_advance(vp(device_ptr), int(count), vp(stream))
    if rc != 0:
        raise TvtError(f"tvt_step_counter_advance failed with status {rc}: {last_error()}")
    launches += 1


def profile_step(fn):
    """Run ``fn`` once with CUDA events around every kernel-launching entry point (on torch's current
    stream, where the kernels are launched) and return {entry point: {ms, flops, calls}}."""
    global _profile
    import torch
    _profile = []
    try:
        fn()
        torch.cuda.synchronize()
        out = {}
        last_profile_detail.clear()
        for name, e0, e1, fl, det in _profile:
            ms = e0.elapsed_time(e1)
            for d in (out.setdefault(name, {"ms": 0.0, "flops": 0.0, "calls": 0}),
                      last_profile_detail.setdefault((name, det), {"ms": 0.0, "flops": 0.0, "calls": 0})):
                d["ms"] += ms
                d["flops"] += fl
                d["calls"] += 1
        return out
    finally:
        _profile = None
