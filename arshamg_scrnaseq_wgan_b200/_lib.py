"""ctypes binding of the C ABI declared in include/wgan_b200.h.

The shared library is built in-tree by ``build()`` (nvcc, sm_100a only) and loaded from the
package directory.  There is no CPU fallback: if the library is missing, or no B200 is
visible when a handle is created, the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, os.environ.get("WGAN_B200_LIB", "libwgan_b200.so"))   # env: A/B builds only
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class WganConfig(C.Structure):
    _fields_ = [
        ("noise_input_size", C.c_int32), ("inflate_to_size", C.c_int32), ("gex_size", C.c_int32),
        ("disc_internal_size", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32),
        ("gemm_mode", C.c_int32), ("gp_form", C.c_int32),
        ("learn_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("epsilon", C.c_float),
        ("alpha", C.c_float), ("lambda_gp", C.c_float), ("dp_rank", C.c_int32), ("dp_world", C.c_int32),
    ]


class WganTensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("net", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32),
                ("ld", C.c_int32), ("offset", C.c_int64)]


# name -> (restype, argtypes); one entry per symbol declared in include/wgan_b200.h
_P, _I, _F = C.c_void_p, C.c_int, C.c_float
SIGNATURES = {
    "wgan_create": (_I, [C.POINTER(WganConfig), C.POINTER(_P)]),
    "wgan_destroy": (None, [_P]),
    "wgan_last_error": (C.c_char_p, [_P]),
    "wgan_tensor_count": (_I, []),
    "wgan_get_tensor_info": (_I, [_P, _I, C.POINTER(WganTensorInfo)]),
    "wgan_get_tensor": (_I, [_P, _I, _I, _P]),
    "wgan_set_tensor": (_I, [_P, _I, _I, _P]),
    "wgan_get_adam_powers": (_I, [_P, _I, C.POINTER(_F), C.POINTER(_F)]),
    "wgan_set_adam_powers": (_I, [_P, _I, _F, _F]),
    "wgan_get_arena": (_I, [_P, _I, _I, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "wgan_get_scalars": (_I, [_P, _I, _P, _P]),
    "wgan_critic_grads": (_I, [_P, _P, _I, _P, _I, _P, _I, _I, C.c_uint64, _P]),
    "wgan_critic_prepare": (_I, [_P, _P, _I, _P, _I, _P, _I, _P, C.POINTER(C.c_uint64)]),
    "wgan_critic_apply": (_I, [_P, _P]),
    "wgan_generator_grads": (_I, [_P, _P, _I, _I, _I, C.c_uint64, _P]),
    "wgan_generator_prepare": (_I, [_P, _P, _I, _I, _P, C.POINTER(C.c_uint64)]),
    "wgan_generator_apply": (_I, [_P, _P]),
    "wgan_dp_export": (_I, [_P, _P]),
    "wgan_dp_connect": (_I, [_P, _P]),
    "wgan_allreduce_grads": (_I, [_P, _I, _P]),
    "wgan_generator_grads_exchange": (_I, [_P, _P, _I, _I, _I, C.c_uint64, _P]),
    "wgan_sample": (_I, [_P, _P, _I, _P, _I, _I, _P]),
    "wgan_critic_forward": (_I, [_P, _P, _I, _P, _I, _P]),
    "wgan_latent_fit_step": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _I, _F, _P]),
    "wgan_rng_noise_prior": (_I, [_P, _I, C.c_int64, _I, _I, _F, C.c_uint64, C.c_uint32, _P, _P]),
    "wgan_rng_uniform": (_I, [_P, C.c_int64, _I, C.c_uint64, C.c_uint32, _P, _P]),
    "wgan_rng_indices": (_I, [_P, C.c_int64, _I, C.c_uint32, C.c_uint64, C.c_uint32, _P, _P]),
    "wgan_counter_add": (_I, [_P, C.c_uint32, _P]),
    "wgan_gather_rows": (_I, [_P, _I, _P, _P, _I, _I, _I, _P]),
    "wgan_draw_minibatch": (_I, [_P, _I, C.c_uint32, _P, _I, _I, _P, _I, _I, _F, _P, _P, C.c_int64, _I, C.c_uint64,
                                 C.c_uint32, _P, _P]),
    "wgan_gemm": (_I, [_P, _I, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P,
                       C.POINTER(_I), _I, _P]),
    "wgan_get_buffer": (_I, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "wgan_launch_count": (C.c_int64, [_P]),
    "wgan_padded_ld": (_I, [_I]),
}


def build(verbose: bool = False) -> str:
    """Compile csrc/engine.cu into libwgan_b200.so for sm_100a (cross-compiles without a GPU)."""
    src = os.path.join(CSRC_DIR, "engine.cu")
    deps = [src] + [os.path.join(CSRC_DIR, f) for f in ("gemm_sm100.cuh", "kernels.cuh", "mid_chain_sm100.cuh")] + \
           [os.path.join(INCLUDE_DIR, "wgan_b200.h")]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the WGAN hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


DP_MAX_WORLD = 16          # WGAN_DP_MAX_WORLD
DP_BLOB_BYTES = 192        # WGAN_DP_BLOB_BYTES


class WganError(RuntimeError):
    pass


def check(lib, handle, rc, what):
    if rc != 0:
        msg = lib.wgan_last_error(handle)
        raise WganError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
