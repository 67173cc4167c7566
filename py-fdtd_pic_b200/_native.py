"""ctypes binding of libpyfdtd_b200.so (the C-ABI declared in include/pyfdtd_b200.h).

There is no CPU fallback: if the shared library is missing, or it reports no usable CUDA device
when a compute entry point is called, this module raises.  PyTorch appears here only as the
carrier of device buffers (``tensor.data_ptr()``) and of the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYFDTD_B200_LIB") or os.path.join(HERE, "libpyfdtd_b200.so")  # env override: kernel tuning builds

PF_FREE, PF_LORENTZ, PF_NL, PF_LORENTZ_NL = 0, 1, 2, 3
PF_ENGINE_OPS, PF_ENGINE_TILE = 0, 1
PF_PIC_F_OFFSETS_VALID = 1
PF_BLOCK_F_TABLES_VALID, PF_BLOCK_F_SWAPPED, PF_BLOCK_F_EDGE_TILES, PF_BLOCK_F_INNER_TILES = 1, 2, 4, 8
PF_ABI_VERSION = 2
PF_F_TFSF, PF_F_CPML_M, PF_F_CPML_P, PF_F_CANONICAL, PF_F_FMA, PF_F_FP32, PF_F_NEWTON = 1, 2, 4, 8, 16, 32, 64

_dp = c_void_p  # device pointers travel as integers


class PfGrid(ctypes.Structure):
    """Mirror of ``struct PfGrid`` (include/pyfdtd_b200.h)."""
    _fields_ = [
        ("L", c_int32), ("pw", c_int32), ("mf", c_int32), ("mr", c_int32), ("nzsrc", c_int32),
        ("flags", c_int32), ("n_probes", c_int32), ("probe_stride", c_int32), ("n_src", c_int32), ("reserved0", c_int32),
        ("z0", c_int64), ("Lg", c_int64),
        ("dt_over_dz", c_double), ("eps0", c_double),
        ("polA", c_double), ("polB", c_double), ("polC", c_double),
        ("cub_a", c_double), ("cub_b", c_double), ("cub_c", c_double),
        ("nl_den0", c_double), ("nl_den1", c_double),
        ("cE0", c_double), ("cE1", c_double), ("cH0", c_double), ("cH1", c_double), ("c2_pml", c_double),
        ("Ex", _dp), ("Hy", _dp), ("Dx", _dp), ("P", _dp), ("Pprev", _dp), ("psiE", _dp), ("psiH", _dp),
        ("Acubic", _dp),
        ("Jx", _dp), ("UpExMat", _dp), ("denE", _dp), ("UpHySelf", _dp), ("UpHyMat", _dp), ("denH", _dp),
        ("beX", _dp), ("ceX", _dp), ("Cb", _dp), ("bmY", _dp), ("cmY", _dp), ("C2", _dp),
        ("srcE", _dp), ("srcH", _dp),
        ("probe_idx", _dp), ("probe_out", _dp),
    ]


class PfPic(ctypes.Structure):
    """Mirror of ``struct PfPic``."""
    _fields_ = [
        ("n", c_int64), ("L", c_int32), ("flags", c_int32),
        ("dz", c_double), ("dt", c_double), ("q_over_m", c_double), ("c", c_double), ("mu0", c_double),
        ("jx_scale", c_double),
        ("z", _dp), ("ux", _dp), ("uz", _dp), ("w", _dp), ("cell", _dp),
        ("z_alt", _dp), ("ux_alt", _dp), ("uz_alt", _dp), ("w_alt", _dp), ("cell_alt", _dp),
        ("Ex", _dp), ("Hy", _dp), ("Jx", _dp),
    ]


class PfDormant(ctypes.Structure):
    """Mirror of ``struct PfDormant``."""
    _fields_ = [
        ("L", c_int32), ("mf", c_int32), ("mr", c_int32), ("reserved0", c_int32),
        ("eps0", c_double), ("dt", c_double),
        ("chi1", c_double), ("chi3", c_double), ("alpha3", c_double), ("one_minus_alpha3", c_double),
        ("lin_AoverD", c_double), ("lin_BoverD", c_double), ("ram_eoverf", c_double), ("ram_hoverf", c_double),
        ("kerr_coef", c_double), ("mur_mult", c_double),
        ("Ex", _dp), ("Eold", _dp), ("Jx", _dp), ("P", _dp), ("Pbar3", _dp), ("Qx3", _dp), ("Gx3", _dp), ("JxKerr", _dp),
    ]


class PfDrudeJ(ctypes.Structure):
    """Mirror of ``struct PfDrudeJ``."""
    _fields_ = [
        ("n", c_int32), ("src", c_int32), ("mat_front", c_int32), ("mat_rear", c_int32), ("n_src", c_int32), ("reserved0", c_int32),
        ("inv_cour", c_double), ("kapE", c_double), ("betaE", c_double), ("c_self", c_double), ("c_curl", c_double),
        ("half_one_plus_kap", c_double),
        ("Ex", _dp), ("Hy", _dp), ("Jx", _dp), ("tempE", _dp), ("tempEOld", _dp), ("Hys", _dp),
    ]


class PfSetupMember(ctypes.Structure):
    """Mirror of ``struct PfSetupMember``."""
    _fields_ = [
        ("L", c_int32), ("pw", c_int32), ("n_src", c_int32), ("src_kind", c_int32), ("tfsf", c_int32), ("pump", c_int32),
        ("dz", c_double), ("delT", c_double), ("eps0", c_double),
        ("kappaMax", c_double), ("r_scale", c_double), ("r_a_scale", c_double), ("sigmaOpt", c_double), ("alphaMax", c_double),
        ("c0", c_double), ("freq", c_double), ("courantNo", c_double), ("period", c_double), ("periods", c_double),
        ("charImp", c_double), ("amp", c_double),
        ("off_beX", c_int64), ("off_ceX", c_int64), ("off_cmY", c_int64), ("off_srcE", c_int64), ("off_srcH", c_int64),
    ]


# every symbol include/pyfdtd_b200.h declares: name -> (restype, argtypes)
_G = POINTER(PfGrid)
_PP = POINTER(PfPic)
SYMBOLS = {
    "pf_abi_version": (c_int, []),
    "pf_last_error": (ctypes.c_char_p, []),
    "pf_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t), POINTER(c_size_t)]),
    "pf_sync": (c_int, [c_void_p]),
    "pf_launch_count": (ctypes.c_ulonglong, []),
    "pf_host_exp": (c_int, [c_void_p, c_void_p, ctypes.c_longlong]),
    "pf_host_pow": (c_int, [c_void_p, c_double, c_void_p, ctypes.c_longlong]),
    "pf_host_sin": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int]),
    "pf_host_cos": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int]),
    "pf_host_sweep_inputs": (c_int, [c_void_p, c_int, c_void_p, c_int]),
    "pf_ade_ex_update": (c_int, [_G, c_void_p]),
    "pf_ade_hy_update": (c_int, [_G, c_void_p]),
    "pf_cpml_psi_e_update": (c_int, [_G, c_void_p]),
    "pf_cpml_psi_m_update": (c_int, [_G, c_void_p]),
    "pf_source_inject": (c_int, [_G, c_int, c_void_p]),
    "pf_ade_dx_update": (c_int, [_G, c_void_p]),
    "pf_ade_polarisation_update": (c_int, [_G, c_void_p]),
    "pf_ade_ex_create": (c_int, [_G, c_void_p]),
    "pf_acubic_finder": (c_int, [_G, c_void_p]),
    "pf_nonlin_ex_update": (c_int, [_G, c_void_p]),
    "pf_probe_record": (c_int, [_G, c_int, c_void_p]),
    "pf_cubic_root0": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "pf_cubic_root0_newton": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "pf_cubic_solve": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pf_varin_pbar": (c_int, [POINTER(PfDormant), c_void_p]),
    "pf_varin_lin_curr_pol": (c_int, [POINTER(PfDormant), c_void_p]),
    "pf_varin_q_and_g": (c_int, [POINTER(PfDormant), c_void_p]),
    "pf_kerr_nonlin": (c_int, [POINTER(PfDormant), c_void_p]),
    "pf_mur1d_ex": (c_int, [POINTER(PfDormant), c_void_p]),
    "pf_drude_j_run": (c_int, [POINTER(PfDrudeJ), c_int, c_int, c_void_p]),
    "pf_run_pass": (c_int, [_G, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pf_run_batch": (c_int, [_G, c_int, c_int, c_int, c_int, POINTER(c_int), c_int, c_void_p, c_size_t, c_void_p]),
    "pf_run_scratch_bytes": (c_size_t, [_G, c_int, c_int]),
    "pf_run_block": (c_int, [_G, _G, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "pf_run_block_scratch_bytes": (c_size_t, [_G, c_int, c_int]),
    "pf_profile_enable": (c_int, [c_int]),
    "pf_profile_collect": (c_int, [POINTER(c_double), POINTER(c_int)]),
    "pf_profile_report": (c_int, [ctypes.c_char_p, c_size_t]),
    "pf_probe_fp64": (c_int, [POINTER(c_double), POINTER(c_double), c_void_p]),
    "pf_tile_config": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "pf_halo_pack": (ctypes.c_longlong, [_G, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pf_halo_unpack": (ctypes.c_longlong, [_G, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pf_pic_push": (c_int, [_PP, c_void_p]),
    "pf_pic_sort": (c_int, [_PP, c_void_p, c_size_t, c_void_p]),
    "pf_pic_push_sorted": (c_int, [_PP, c_void_p, c_size_t, c_void_p]),
    "pf_pic_check": (c_int, [_PP, c_void_p, c_size_t, c_void_p]),
    "pf_pic_deposit": (c_int, [_PP, c_void_p, c_size_t, c_void_p]),
    "pf_pic_step_sorted": (c_int, [_PP, c_void_p, c_size_t, c_void_p]),
    "pf_pic_sub_warps": (c_int, [_PP]),
    "pf_pic_scratch_bytes": (c_size_t, [_PP]),
}


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library (once).  Raises NativeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the hot path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if L.pf_abi_version() != PF_ABI_VERSION:
            raise NativeError("libpyfdtd_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc < 0:
        msg = lib().pf_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise NativeError(f"{what}: {msg} (code {rc})")
    return rc


def require_cuda():
    """Fail loudly when the hot path is asked to run without a GPU."""
    import torch
    if not torch.cuda.is_available():
        raise NativeError("pyfdtd_b200 needs a CUDA device (sm_100a); no CPU fallback exists for the hot path")
    return torch


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def profile_report():
    """{kernel name: (launches, total ms)} of the launches timed since pf_profile_enable(1) / the last report."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().pf_profile_report(buf, len(buf)), "pf_profile_report")
    out = {}
    for item in buf.value.decode().split(";"):
        if item:
            name, n, ms = item.rsplit("|", 2)
            out[name] = (int(n), float(ms))
    return out


def probe_fp64():
    """(separately rounded DMUL+DADD instructions/s, DFMA/s) of the current device, measured now."""
    a, b = c_double(), c_double()
    check(lib().pf_probe_fp64(a, b, current_stream_ptr()), "pf_probe_fp64")
    return a.value, b.value


def device_info():
    sm, ma, mi = c_int(), c_int(), c_int()
    fr, to = c_size_t(), c_size_t()
    check(lib().pf_device_info(sm, ma, mi, fr, to), "pf_device_info")
    return dict(sm_count=sm.value, cc=(ma.value, mi.value), free_bytes=fr.value, total_bytes=to.value)
