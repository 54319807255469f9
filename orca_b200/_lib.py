"""ctypes binding of liborca_b200.so (the C ABI declared in include/orca_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m orca_b200.build``.
If it is missing the import of any compute entry point fails loudly -- there is no fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborca_b200.so")

OK = 0
ENCODER, ENCODER2, ENCODER2B, ENCODER3, DECODER, DECODER_1M, NET = 1, 2, 3, 4, 5, 6, 7
UPSAMPLE_NEAREST, UPSAMPLE_BILINEAR = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
OPT_IMPL, OPT_ENCODER_FP16_STAGES = 1, 2
DEFAULT_FP16_STAGES = 4  # kDefaultFp16Stages in csrc/modules.cu
STATUS_FP16_RANGE = 1

_fp = ctypes.POINTER(ctypes.c_float)
_i64 = ctypes.c_int64
_vp = ctypes.c_void_p


class Region(ctypes.Structure):
    """struct orca_b200_region"""
    _fields_ = [("chrom", ctypes.c_int32), ("reverse", ctypes.c_int32), ("start", ctypes.c_int64), ("end", ctypes.c_int64)]


class ConvParams(ctypes.Structure):
    """struct orca_b200_conv_params"""
    _fields_ = [("c_in", ctypes.c_int32), ("c_out", ctypes.c_int32), ("kh", ctypes.c_int32),
                ("kw", ctypes.c_int32), ("dilation", ctypes.c_int32),
                ("weight", _fp), ("bias", _fp), ("bn_weight", _fp), ("bn_bias", _fp),
                ("bn_mean", _fp), ("bn_var", _fp), ("bn_eps", ctypes.c_float)]


# name -> (restype, argtypes); mirrors include/orca_b200.h one to one
SIGNATURES = {
    "orca_b200_version": (ctypes.c_char_p, []),
    "orca_b200_last_error": (ctypes.c_char_p, []),
    "orca_b200_launch_count": (ctypes.c_uint64, []),
    "orca_b200_module_set_option": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    "orca_b200_module_get_option": (ctypes.c_int, [_vp, ctypes.c_int]),
    "orca_b200_module_status": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_uint32), ctypes.c_int32]),
    "orca_b200_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "orca_b200_profile_summary": (ctypes.c_int64, [ctypes.c_char_p, ctypes.c_int64]),
    "orca_b200_module_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ConvParams), ctypes.c_int32,
                                               ctypes.c_uint32, ctypes.c_int32, ctypes.POINTER(_vp)]),
    "orca_b200_module_destroy": (None, [_vp]),
    "orca_b200_module_kind": (ctypes.c_int, [_vp]),
    "orca_b200_module_num_2d": (ctypes.c_int, [_vp]),
    "orca_b200_encoder_workspace_bytes": (ctypes.c_size_t, [_vp, _i64, _i64, _i64]),
    "orca_b200_encoder_forward": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64,
                                                 _i64, _vp, ctypes.c_size_t, _vp]),
    "orca_b200_encoder_forward_packed": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, ctypes.c_int32, _i64, _i64, _vp,
                                                        _i64, _i64, _i64, _vp, ctypes.c_size_t, _vp]),
    "orca_b200_encoder2_workspace_bytes": (ctypes.c_size_t, [_vp, _i64, _i64]),
    "orca_b200_encoder2_forward": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, ctypes.POINTER(_vp),
                                                  ctypes.c_int32, ctypes.c_int32, _vp, ctypes.c_size_t, _vp]),
    "orca_b200_decoder_workspace_bytes": (ctypes.c_size_t, [_vp, _i64, _i64]),
    "orca_b200_decoder_forward": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64,
                                                 _vp, _i64, _i64, _i64, _i64, _vp, _vp, ctypes.c_size_t, _vp]),
    "orca_b200_net_workspace_bytes": (ctypes.c_size_t, [_vp, _i64, _i64]),
    "orca_b200_net_forward": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp,
                                             ctypes.c_size_t, _vp]),
    "orca_b200_net_forward_packed": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, ctypes.c_int32, _vp, _vp, _vp,
                                                    ctypes.c_size_t, _vp]),
    "orca_b200_background_forward": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, ctypes.c_int32, _vp, _vp]),
    "orca_b200_background_level": (ctypes.c_int, [_vp, _i64, _i64, _i64, _i64, ctypes.c_int32, _vp, _vp, _vp]),
    "orca_b200_background_bins": (ctypes.c_int64, [ctypes.POINTER(Region), ctypes.c_int32, _i64]),
    "orca_b200_background_assemble": (ctypes.c_int, [ctypes.POINTER(Region), ctypes.c_int32, _vp, _i64, ctypes.c_double, _i64,
                                                     _vp, _i64, _vp, ctypes.c_size_t, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "orca_b200: %s is missing -- build it with `python -m orca_b200.build` "
                "(there is no CPU or PyTorch fallback for the compute path)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status):
    if status != OK:
        msg = lib().orca_b200_last_error()
        raise RuntimeError("orca_b200 error %d: %s" % (status, (msg or b"?").decode("utf-8", "replace")))


# Python-side defaults applied to module handles (orca_b200.modules._NativeModule.native_handle): the C library itself
# keeps kernel selection and precision per handle.  `options_epoch` changes whenever a default changes, so existing
# modules re-apply the options before their next forward.
_defaults = {"impl": IMPL_AUTO, "encoder_fp16_stages": -1}
options_epoch = 0


def set_impl(impl):
    """Default kernel family for every module: "auto" (tcgen05 where available), "simt" (fp32 CUDA cores), "tc"."""
    global options_epoch
    v = {"auto": IMPL_AUTO, "simt": IMPL_SIMT, "tc": IMPL_TC}.get(impl, impl)
    if v not in (IMPL_AUTO, IMPL_SIMT, IMPL_TC):
        raise ValueError("unknown impl %r" % (impl,))
    _defaults["impl"] = v
    options_epoch += 1


def set_encoder_fp16_stages(n):
    """Default number of leading encoder stages in single-pass fp16 (0..7, -1 = library default 4); returns the
    previous effective setting."""
    global options_epoch
    prev = _defaults["encoder_fp16_stages"]
    _defaults["encoder_fp16_stages"] = -1 if n < 0 else min(int(n), 7)
    options_epoch += 1
    return DEFAULT_FP16_STAGES if prev < 0 else prev


def apply_options(handle, overrides=None):
    """Push the defaults (plus a module's own overrides) into a native handle."""
    opts = dict(_defaults)
    if overrides:
        opts.update(overrides)
    check(lib().orca_b200_module_set_option(handle, OPT_IMPL, opts["impl"]))
    check(lib().orca_b200_module_set_option(handle, OPT_ENCODER_FP16_STAGES, opts["encoder_fp16_stages"]))


def module_status(handle, clear=True):
    """Device status word of a handle (synchronises): bit 0 = the fp16 range guard fired."""
    st = ctypes.c_uint32(0)
    check(lib().orca_b200_module_status(handle, ctypes.byref(st), 1 if clear else 0))
    return int(st.value)


def launch_count():
    return int(lib().orca_b200_launch_count())


def profile_enable(on):
    check(lib().orca_b200_profile_enable(1 if on else 0))


def profile_summary():
    """Aggregated per-shape conv timings recorded since profile_enable(True) (synchronises)."""
    import json
    n = lib().orca_b200_profile_summary(None, 0)
    if n < 0:
        check(int(n))
    buf = ctypes.create_string_buffer(int(n) + 16)
    lib().orca_b200_profile_summary(buf, len(buf))
    return json.loads(buf.value.decode())
