"""ctypes binding of librcn_b200.so (the C ABI declared in include/rcn_b200.h).

There is NO fallback: if the shared library is missing or fails to load, importing any op
raises.  Build it with ``python -m realcamnet_b200.build`` (nvcc, sm_100a).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RCN_B200_LIB") or os.path.join(_HERE, "librcn_b200.so")


class ConvDesc(ctypes.Structure):
    """struct rcn_conv_desc"""
    _fields_ = [
        ("x", c_void_p), ("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("ldx", c_int),
        ("w", c_void_p), ("bias", c_void_p),
        ("k", c_int), ("stride", c_int), ("Cout", c_int), ("in_square", c_int),
        ("y", c_void_p), ("ldy", c_int), ("store", c_int),
        ("epi", c_int), ("aux", c_void_p), ("ldaux", c_int),
        ("cscale", c_void_p), ("cshift", c_void_p),
        ("res", c_void_p), ("ldres", c_int), ("res_pre", c_int),
        ("act", c_int), ("slope", c_float), ("res_scale", c_float),
        ("y_hi", c_void_p), ("y_lo", c_void_p), ("Cp_out", c_int), ("ldp_in", c_int), ("planes_s2", c_int), ("ps_perm", c_int),
        ("aux_nchw", c_int), ("planes_square", c_int), ("in_fmt", c_int), ("out_fmt", c_int),
    ]


class IngestDesc(ctypes.Structure):
    """struct rcn_ingest_desc"""
    _fields_ = [
        ("coord", c_void_p), ("coord_bs", c_longlong), ("coord_ps", c_int), ("coord_cs", c_int),
        ("w0", c_void_p), ("b0", c_void_p),
        ("w1_hi", c_void_p), ("w1_lo", c_void_p), ("w2_hi", c_void_p), ("w2_lo", c_void_p), ("w3_hi", c_void_p), ("w3_lo", c_void_p),
        ("b1", c_void_p), ("b2", c_void_p), ("b3", c_void_p),
        ("slope", c_float),
        ("lsc", c_void_p), ("N", c_int), ("H", c_int), ("W", c_int),
        ("raw", c_void_p), ("ldraw", c_int),
        ("wc_hi", c_void_p), ("wc_lo", c_void_p), ("bc", c_void_p),
        ("fea_hi", c_void_p), ("fea_lo", c_void_p), ("planes_s2", c_int),
    ]


class MlpDesc(ctypes.Structure):
    """struct rcn_mlp_desc"""
    _fields_ = [
        ("x_hi", c_void_p), ("x_lo", c_void_p), ("ldp_in", c_int),
        ("npix", c_longlong), ("C", c_int), ("hidden", c_int),
        ("w1_hi", c_void_p), ("w1_lo", c_void_p), ("b1", c_void_p),
        ("w2_hi", c_void_p), ("w2_lo", c_void_p), ("b2", c_void_p),
        ("res", c_void_p), ("ldres", c_int),
        ("y", c_void_p), ("ldy", c_int),
        ("y_hi", c_void_p), ("y_lo", c_void_p), ("Cp_out", c_int),
        ("x_ln", c_void_p), ("ldx", c_int), ("gamma", c_void_p), ("beta", c_void_p), ("eps", c_float),
    ]


class LnLinearDesc(ctypes.Structure):
    """struct rcn_lnlinear_desc"""
    _fields_ = [
        ("x", c_void_p), ("ldx", c_int), ("npix", c_longlong), ("C", c_int), ("Cout", c_int),
        ("gamma", c_void_p), ("beta", c_void_p), ("eps", c_float),
        ("w_hi", c_void_p), ("w_lo", c_void_p), ("bias", c_void_p),
        ("y", c_void_p), ("ldy", c_int),
    ]


_P, _I, _L, _F = c_void_p, c_int, c_longlong, c_float

# name -> (restype, argtypes); must list every symbol of include/rcn_b200.h (tests check this)
PROTOTYPES = {
    "rcn_last_error": (c_char_p, []),
    "rcn_version": (_I, []),
    "rcn_launch_count": (c_ulonglong, []),
    "rcn_conv2d": (_I, [POINTER(ConvDesc), _P]),
    "rcn_conv2d_tc": (_I, [POINTER(ConvDesc), _P, _P, _P, _P, _I, _I, _P]),
    "rcn_split_bf16": (_I, [_P, _I, _L, _I, _I, _I, _I, _P, _P, _P]),
    "rcn_split_bf16_s2": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "rcn_pack_conv_weight_tc": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "rcn_tc_prof": (_I, [_P, _I]),
    "rcn_ingest_fused": (_I, [POINTER(IngestDesc), _P]),
    "rcn_mlp_fused": (_I, [POINTER(MlpDesc), _P]),
    "rcn_ln_linear_fused": (_I, [POINTER(LnLinearDesc), _P]),
    "rcn_pack_ingest_weight": (_I, [_P, _P, _P, _P]),
    "rcn_pack_conv_weight": (_I, [_P, _I, _I, _I, _P, _P]),
    "rcn_layernorm": (_I, [_P, _L, _I, _I, _P, _P, _F, _P, _I, _I, _P, _P, _I, _P]),
    "rcn_wmsa": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P]),
    "rcn_nchw_to_nhwc": (_I, [_P, _I, _I, _I, _I, _P, _I, _P]),
    "rcn_nhwc_to_nchw": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "rcn_copy_channels": (_I, [_P, _I, _L, _I, _P, _I, _P]),
    "rcn_channel_mean": (_I, [_P, _I, _L, _I, _I, _P, _P, _L, _P]),
    "rcn_channel_meanvar": (_I, [_P, _I, _L, _I, _I, _P, _P, _P]),
    "rcn_norm_apply": (_I, [_P, _I, _I, _L, _I, _P, _P, _P, _P, _F, _P, _I, _P]),
    "rcn_scale_add": (_I, [_P, _I, _I, _L, _I, _P, _P, _I, _P, _I, _P, _I, _I, _P]),
    "rcn_avgpool3s2_lrelu": (_I, [_P, _I, _I, _I, _I, _I, _F, _P, _I, _P]),
    "rcn_upsample_bilinear2x": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "rcn_upsample_bilinear2x_planes": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "rcn_dwt_forward": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "rcn_dwt_inverse": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "rcn_space_to_depth2": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P]),
    "rcn_depthwise_conv": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P, _I, _P]),
    "rcn_eb_forward": (_I, [_P, _I, _I, _L, _I, _P, _P, _P, _I, _P, _I, _P, _F, _P]),
    "rcn_eb_dequantize": (_I, [_P, _I, _L, _I, _P, _P, _I, _P]),
    "rcn_gaussian_conditional": (_I, [_P, _I, _P, _I, _P, _I, _I, _L, _I, _P, _I, _F, _F, _P, _I, _P, _I, _P, _P, _P]),
    "rcn_gaussian_conditional_coded": (_I, [_P, _I, _P, _I, _P, _I, _I, _L, _I, _P, _I, _F, _F, _P, _I, _P, _I, _P, _P,
                                             _P, _I, _P, _P, _P, _P, _P, _P]),
    "rcn_build_indexes": (_I, [_P, _I, _I, _L, _I, _P, _I, _F, _P, _P]),
    "rcn_gaussian_dequantize": (_I, [_P, _P, _I, _I, _L, _I, _P, _I, _P]),
    "rcn_groupmix_workspace_floats": (_L, [_I, _L, _I, _I]),
    "rcn_groupmix_attention": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _L, _I, _I, _F, _P, _I, _P, _P, _L, _P]),
    "rcn_rans_encode": (_L, [_P, _P, _L, _P, _I, _P, _P, _P, _L]),
    "rcn_rans_encode_packed": (_L, [_P, _P, _P, _L, _P, _L]),
    "rcn_rans_decoder_create": (_P, [_P, _L]),
    "rcn_rans_decode": (_I, [_P, _P, _L, _P, _I, _I, _P, _P, _P]),
    "rcn_rans_decoder_destroy": (None, [_P]),
    "rcn_pmf_to_quantized_cdf": (_I, [_P, _I, _I, _P]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"realcamnet_b200: {LIB_PATH} is missing -- build the CUDA library with "
                "`python -m realcamnet_b200.build` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc is None or rc >= 0:
        return rc
    msg = lib().rcn_last_error()
    raise RuntimeError(f"{what or 'librcn_b200'} failed ({rc}): {msg.decode() if msg else ''}")
