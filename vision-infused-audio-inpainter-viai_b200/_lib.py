"""ctypes binding of libviai_b200.so (the C ABI declared in include/viai_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libviai_b200.so")
_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_l = ctypes.c_int64
c_f = ctypes.c_float
c_d = ctypes.c_double


class ConvGeom(ctypes.Structure):
    """Mirror of ``viai_conv_geom``."""
    _fields_ = [(n, ctypes.c_int32) for n in (
        "N", "Hin", "Win", "Cin", "Hout", "Wout", "Cout", "R", "S", "stride_h", "stride_w", "pad_h", "pad_w", "mode")]


class NormBwdCtx(ctypes.Structure):
    """Mirror of ``viai_norm_bwd_ctx``."""
    _fields_ = [("y", ctypes.c_void_p), ("mean", ctypes.c_void_p), ("invstd", ctypes.c_void_p), ("gamma", ctypes.c_void_p),
                ("beta", ctypes.c_void_p), ("act", ctypes.c_int32), ("slope", ctypes.c_float)]


class PackDesc(ctypes.Structure):
    """Mirror of ``viai_pack_desc``."""
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("O", ctypes.c_int32), ("I", ctypes.c_int32), ("R", ctypes.c_int32),
                ("S", ctypes.c_int32), ("so", ctypes.c_int64), ("si", ctypes.c_int64), ("sr", ctypes.c_int64), ("ss", ctypes.c_int64),
                ("flip", ctypes.c_int32), ("kind", ctypes.c_int32)]


_GP = ctypes.POINTER(ConvGeom)

# name -> argtypes (return type is always int)
SIGNATURES = {
    "viai_pack_weight": [c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_l, c_l, c_l, c_i, c_p],
    "viai_conv2d_simt": [_GP, c_p, c_p, c_p, c_p, c_p],
    "viai_pack_weight_tc": [c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_l, c_l, c_l, c_i, c_i, c_p],
    "viai_pack_weights_batched": [ctypes.POINTER(PackDesc), c_i, c_p],
    "viai_conv2d_tc_supported": [_GP],
    "viai_conv2d_tc": [_GP, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p],
    "viai_conv2d_tc_bwd_reduce": [_GP, c_p, c_p, c_p, ctypes.POINTER(NormBwdCtx), c_p, c_p, c_i, c_p],
    "viai_tc_bn": [c_i],
    "viai_tc_f16_overflow": [c_i, ctypes.POINTER(ctypes.c_uint)],
    "viai_conv2d_wgrad_tc_supported": [_GP],
    "viai_conv2d_wgrad_tc": [_GP, c_p, c_p, c_p, c_l, c_l, c_l, c_l, c_i, c_p, c_p],
    "viai_conv2d_thin_supported": [_GP],
    "viai_conv2d_thin": [_GP, c_p, c_p, c_p, c_p, c_p],
    "viai_conv2d_thin_stats_supported": [_GP],
    "viai_conv2d_thin_stats": [_GP, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "viai_conv2d_wgrad_thin_supported": [_GP],
    "viai_conv2d_wgrad_thin": [_GP, c_p, c_p, c_p, c_l, c_l, c_l, c_l, c_i, c_p, c_p],
    "viai_conv2d_wgrad_simt": [_GP, c_p, c_p, c_p, c_l, c_l, c_l, c_l, c_i, c_p],
    "viai_norm_walk_mb": [c_i],
    "viai_channel_stats": [c_p, c_l, c_i, c_i, c_p, c_p, c_p],
    "viai_norm_finalize": [c_p, c_p, c_l, c_i, c_i, c_f, c_p, c_p, c_p, c_p, c_f, c_p, c_p],
    "viai_norm_act_fwd": [c_p, c_l, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_f, c_p, c_p],
    "viai_norm_act_bwd_reduce": [c_p, c_p, c_l, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_f, c_p, c_p, c_p],
    "viai_norm_act_bwd_apply": [c_p, c_p, c_l, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_p],
    "viai_norm_act_bwd_apply_fold": [c_p, c_p, c_l, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_i, c_p],
    "viai_norm_finalize_act_fwd": [c_p, c_l, c_i, c_i, c_p, c_p, c_f, c_p, c_p, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p],
    "viai_fold_groups": [c_p, c_i, c_i, c_p, c_i, c_p],
    "viai_rsqrt_eps": [c_p, c_i, c_f, c_p, c_p],
    "viai_bilinear_fwd": [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_p],
    "viai_bilinear_bwd": [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_i, c_p],
    "viai_copy_channels": [c_p, c_l, c_i, c_i, c_p, c_i, c_i, c_i, c_p],
    "viai_avgpool_h_fwd": [c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "viai_avgpool_h_bwd": [c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "viai_maxpool3s2_fwd": [c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p],
    "viai_maxpool3s2_bwd": [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p],
    "viai_maxpool3s2_bwd_out": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p],
    "viai_maxpool3s2_fwd_idx": [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_p],
    "viai_maxpool3s2_bwd_idx": [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_i, c_i, c_p],
    "viai_im2col": [_GP, c_p, c_i, c_p, c_p],
    "viai_mul": [c_p, c_p, c_p, c_l, c_p],
    "viai_add_act": [c_p, c_p, c_p, c_l, c_i, c_p],
    "viai_add_act_bwd": [c_p, c_p, c_p, c_l, c_i, c_p],
    "viai_stft_mel": [c_p, c_l, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_i, c_f, c_f, c_p, c_p, c_p],
    "viai_wavenet_num_ctas": [c_i] * 7,
    "viai_wavenet_synth": [c_i] * 11 + [c_p] * 7 + [c_i, c_f] + [c_p] * 9,
    "viai_wavenet3_num_ctas": [c_i] * 8,
    "viai_wavenet3_replicas": [],
    "viai_wavenet3_profile": [c_p],
    "viai_wavenet_synth3": [c_i] * 11 + [c_p] * 8 + [c_i, c_f] + [c_p] * 9,
    "viai_wavenet2_num_ctas": [c_i] * 8,
    "viai_wavenet2_profile": [c_p],
    "viai_wavenet_synth2": [c_i] * 11 + [c_p] * 8 + [c_i, c_f] + [c_p] * 9,
    "viai_wavenet_cluster_supported": [c_i] * 7,
    "viai_wavenet_synth_cluster": [c_i] * 10 + [c_p] * 7 + [c_i, c_f] + [c_p] * 5,
    "viai_loss_fwd": [c_i, c_p, c_p, c_f, c_l, c_p, c_p, c_p],
    "viai_loss_bwd": [c_i, c_p, c_p, c_f, c_l, c_p, c_p, c_p],
    "viai_adam_step": [c_p, c_p, c_p, c_p, c_l, c_p, c_d, c_d, c_d, c_p, c_i, c_f, c_p],
    "viai_lincomb2": [c_p, c_f, c_p, c_f, c_p, c_p],
    "viai_fill": [c_p, c_l, c_f, c_p],
    "viai_shiftcat_fwd": [c_p, c_p, c_p, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p],
    "viai_shiftcat_bwd": [c_p, c_p, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p],
    "viai_weight_norm_fwd": [c_p, c_p, c_i, c_i, c_p, c_p, c_p],
    "viai_weight_norm_bwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p],
    "viai_glu_fwd": [c_p, c_l, c_i, c_p, c_p],
    "viai_glu_bwd": [c_p, c_p, c_l, c_i, c_p, c_p],
    "viai_axpby": [c_p, c_f, c_p, c_f, c_p, c_l, c_p],
    "viai_dmol_nll": [c_p, c_p, c_l, c_i, c_i, c_f, c_p, c_p, c_p, c_p],
    "viai_dmol_sample": [c_p, c_p, c_l, c_i, c_f, c_p, c_p],
    "viai_masked_sum_fwd": [c_p, c_p, c_l, c_i, c_p, c_p, c_p],
    "viai_masked_sum_bwd": [c_p, c_l, c_i, c_p, c_p, c_p, c_p],
    "viai_sequence_mask": [c_p, c_i, c_i, c_p, c_p],
    "viai_frames_preprocess": [c_p] + [c_i] * 14 + [c_p, c_p],
    "viai_l2norm_fwd": [c_p, c_i, c_i, c_f, c_p, c_p, c_p],
    "viai_l2norm_bwd": [c_p, c_p, c_p, c_i, c_i, c_f, c_p, c_p],
    "viai_pairdist_fwd": [c_p, c_p, c_i, c_i, c_i, c_p, c_p],
    "viai_pairdist_bwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p],
    "viai_retrieval_ranks": [c_p, c_i, c_i, c_p, c_p, c_p],
    "viai_l2_contrastive": [c_p, c_i, c_f, c_i, c_p, c_p, c_p, c_p],
}


# entry points whose return type is not the int status code: name -> (restype, argtypes)
VALUE_FUNCS = {
    "viai_tc_packed_size": (ctypes.c_int64, [c_i, c_i, c_i, c_i, c_i]),
    "viai_wgrad_tc_workspace": (ctypes.c_int64, [_GP]),
    "viai_wgrad_thin_workspace": (ctypes.c_int64, [_GP]),
}


def available():
    return os.path.exists(LIB_PATH)


def lib():
    """Loads the shared library once.  Raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libviai_b200.so not found at %s -- build it with __graft_entry__.build(); "
                               "there is no CPU or PyTorch fallback for the VIAI hot path" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.viai_last_error.restype = ctypes.c_char_p
        L.viai_last_error.argtypes = []
        L.viai_version.restype = c_i
        L.viai_launch_count.restype = ctypes.c_longlong
        for name, (res, args) in VALUE_FUNCS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = c_i
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().viai_last_error().decode("utf-8", "replace")
        raise RuntimeError("libviai_b200 %s failed (%d): %s" % (what, rc, msg))


def launch_count():
    return int(lib().viai_launch_count())
