"""ctypes binding of libdamc_b200.so (C ABI in include/damc.h).  No CPU fallback: if the library is missing or fails
to load, importing the samplers raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdamc_b200.so")

OK = 0
PREC_FP32, PREC_BF16, PREC_FP16, PREC_TF32 = 0, 1, 2, 3


class ConvTLayer(C.Structure):
    _fields_ = [("cin", C.c_int), ("cout", C.c_int), ("k", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("weight", C.c_void_p), ("bias", C.c_void_p)]


class ConvLayer(C.Structure):
    _fields_ = [("cin", C.c_int), ("cout", C.c_int), ("k", C.c_int), ("stride", C.c_int), ("pad", C.c_int),
                ("weight", C.c_void_p), ("bias", C.c_void_p), ("in_weight", C.c_void_p), ("in_bias", C.c_void_p)]


class DenoiserDesc(C.Structure):
    _fields_ = [("nz", C.c_int), ("nxemb", C.c_int), ("ntemb", C.c_int), ("nf", C.c_int), ("residual", C.c_int),
                ("time_w1", C.c_void_p), ("time_b1", C.c_void_p), ("time_w2", C.c_void_p), ("time_b2", C.c_void_p),
                ("Bproj", C.c_void_p), ("dim_in", C.c_int * 7), ("dim_out", C.c_int * 7),
                ("W", C.c_void_p * 7), ("b", C.c_void_p * 7), ("Wc", C.c_void_p * 7), ("bc", C.c_void_p * 7),
                ("Wg", C.c_void_p * 7), ("bg", C.c_void_p * 7), ("Wb", C.c_void_p * 7), ("Ws", C.c_void_p * 7),
                ("bs", C.c_void_p * 7)]


# symbol -> (restype, argtypes); every symbol include/damc.h declares must appear here (tests check the export list)
_P, _I, _F, _U64, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_size_t
SIGNATURES = {
    "damc_version": (_I, []),
    "damc_last_error": (C.c_char_p, []),
    "damc_free": (_I, [_P]),
    "damc_repack": (_I, [_P, _P]),
    "damc_pack_mlp": (_I, [C.POINTER(_P), _I, _I, _P, _P, _P, _P, _P, _P, _F, _P]),
    "damc_pack_generator": (_I, [C.POINTER(_P), _I, C.POINTER(ConvTLayer), _F, _I, _P]),
    "damc_generator_shape": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "damc_generator_workspace_bytes": (_SZ, [_P, _I]),
    "damc_generator_forward": (_I, [_P, _P, _P, _I, _P, _SZ, _P]),
    "damc_posterior_score": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _SZ, _P]),
    "damc_fused_clip_adam": (_I, [_P, _P, _P, _P, _SZ, _I, _P, _P, _P, _F, _F, _F, _F, _F, _I, _F, _F, _P, _P]),
    "damc_round_tf32": (_I, [_P, _P, _SZ, _P]),
    "damc_gemm_tf32": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "damc_prior_langevin": (_I, [_P, _P, _I, _I, _F, _I, _P, _U64, _U64, _U64, _P, _P]),
    "damc_prior_langevin_tc": (_I, [_P, _P, _I, _I, _F, _I, _P, _U64, _U64, _U64, _P]),
    "damc_posterior_langevin": (_I, [_P, _P, _P, _P, _I, _I, _F, _F, _I, _P, _U64, _U64, _U64, _P, _P, _P, _SZ, _P]),
    "damc_pack_toy_mlp": (_I, [C.POINTER(_P), _I, _I, _I, C.POINTER(_P), C.POINTER(_P), _P]),
    "damc_toy_posterior_langevin": (_I, [_P, _P, _P, _I, _I, _F, _F, _I, _P, _U64, _U64, _U64, _P]),
    "damc_pack_denoiser": (_I, [C.POINTER(_P), C.POINTER(DenoiserDesc), _P]),
    "damc_denoise_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "damc_denoise": (_I, [_P, _P, _P, _I, _I, C.POINTER(_F), _I, _I, _P, _U64, _U64, _I, _P, _SZ, _P]),
    "damc_denoiser_eps": (_I, [_P, _P, _P, _F, _P, _I, _I, _P, _SZ, _P]),
    "damc_pack_encoder": (_I, [C.POINTER(_P), _I, C.POINTER(ConvLayer), _I, _I, _F, _F, _I, _P]),
    "damc_encoder_workspace_bytes": (_SZ, [_P, _I]),
    "damc_encoder_forward": (_I, [_P, _P, _P, _I, _P, _SZ, _P]),
    "damc_selftest": (_I, []),
    "damc_launch_count": (C.c_longlong, []),
    "damc_graph_replays": (C.c_longlong, [_P]),
    "damc_profile_enable": (_I, [_I]),
    "damc_profile_collect": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
}

_lib = None


def lib():
    """Load (once) and return the CDLL.  Raises RuntimeError if the CUDA library is absent -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). damc_b200 has no CPU or PyTorch fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)  # AttributeError if the .so is stale / incomplete
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc, what=""):
    if rc != OK:
        msg = lib().damc_last_error().decode(errors="replace")
        raise RuntimeError(f"libdamc_b200 {what} failed (code {rc}): {msg}")
