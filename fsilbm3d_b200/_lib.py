"""ctypes binding of libfsilbm_b200.so (include/fsilbm.h).  Fails loudly: no fallback of any kind."""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "fsilbm.h")


class FsilbmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fsilbm error {code}: {msg}")
        self.code = code


class CFlow(C.Structure):
    _fields_ = [
        ("nu", C.c_double), ("denIn", C.c_double),
        ("uvwIn", C.c_double * 3), ("shearRateIn", C.c_double * 3),
        ("velocityKind", C.c_int),
        ("volumeForceIn", C.c_double * 3), ("volumeForceAmp", C.c_double),
        ("volumeForceFreq", C.c_double), ("volumeForcePhi", C.c_double),
        ("Uref", C.c_double),
    ]


def library_path() -> str:
    return os.path.join(HERE, "libfsilbm_b200.so")


def declared_symbols() -> List[str]:
    """Every function include/fsilbm.h declares."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fsilbm_[a-zA-Z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    """Load the CUDA library (building it first if the .so is missing and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        from .build import build
        build()
    L = C.CDLL(path)
    i, d, vp = C.c_int, C.c_double, C.c_void_p
    pi, pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
    ppd = C.POINTER(C.c_void_p)
    sig = {
        "fsilbm_init": [i], "fsilbm_finalize": [], "fsilbm_set_option": [C.c_char_p, i], "fsilbm_trace_dump": [C.c_char_p],
        "fsilbm_block_create": [i, i, i, i, i, d, d, d, d, pi, i, pd, C.POINTER(CFlow), pi],
        "fsilbm_block_destroy": [i], "fsilbm_block_initialise": [i, d], "fsilbm_block_get": [i, i, pd],
        "fsilbm_block_upload_fIn": [i, vp], "fsilbm_block_download_fIn": [i, vp],
        "fsilbm_block_set_time": [i, d], "fsilbm_block_update_volume_force": [i, pd],
        "fsilbm_block_download_macro": [i, vp, vp],
        "fsilbm_block_download_macro_async": [i, vp, vp],
        "fsilbm_block_download_wait": [i], "fsilbm_block_download_tau_all": [i, vp], "fsilbm_block_field_stat": [i, pd],
        "fsilbm_block_write_flow_window": [i, i, i, vp], "fsilbm_block_write_flow_window_async": [i, i, i, vp],
        "fsilbm_block_turbulent_statistic": [i, i, i],
        "fsilbm_block_fluid_flux": [i, pd], "fsilbm_block_probe_velocity": [i, i, vp, vp],
        "fsilbm_block_set_boundary_conditions": [i], "fsilbm_block_collide_stream": [i], "fsilbm_block_sync": [i],
        "fsilbm_block_stream": [i, C.POINTER(C.c_void_p)],
        "fsilbm_block_halo_transport": [i, pi],
        "fsilbm_block_pass_macro": [i], "fsilbm_block_pass_reset_volume_force": [i],
        "fsilbm_block_pass_add_volume_force": [i], "fsilbm_block_pass_collision": [i],
        "fsilbm_block_pass_halfway_bc_set": [i], "fsilbm_block_pass_streaming": [i],
        "fsilbm_block_download_fields": [i, vp, vp, vp], "fsilbm_block_upload_fields": [i, vp, vp, vp],
        "fsilbm_ibm_interaction_force": [i, i, pi, ppd, ppd, ppd, ppd, pi, d, i, d, pi, pi],
        "fsilbm_ibm_interaction_force_begin": [i, i, pi, ppd, ppd, ppd, pi, d, i, d, pi],
        "fsilbm_ibm_interaction_force_wait": [i, i, ppd, pi],
        "fsilbm_ibm_download_stencil": [i, i, vp, vp],
        "fsilbm_ibm_body_status": [i, i, vp],
        "fsilbm_pair_create": [i, i, i, pi], "fsilbm_pair_create_remote": [i, i, pi], "fsilbm_pair_destroy": [i], "fsilbm_pair_info": [i, pi],
        "fsilbm_pair_extract_layer": [i, i], "fsilbm_pair_father_to_son": [i, i], "fsilbm_pair_son_to_father": [i],
        "fsilbm_comm_unique_id": [C.c_char_p], "fsilbm_comm_init": [i, i, C.c_char_p], "fsilbm_comm_finalize": [],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = i
    L.fsilbm_last_error.restype = C.c_char_p
    L.fsilbm_last_error.argtypes = []
    L.fsilbm_launch_count.restype = C.c_longlong
    L.fsilbm_launch_count.argtypes = []
    L.fsilbm_ibm_early_count.restype = C.c_longlong
    L.fsilbm_ibm_early_count.argtypes = []
    _lib = L
    return L


def exported_symbols() -> List[str]:
    """Names from the header that the built library actually exports (no compute call made)."""
    L = C.CDLL(library_path())
    return [s for s in declared_symbols() if hasattr(L, s)]


def check(rc: int) -> None:
    if rc != 0:
        raise FsilbmError(rc, lib().fsilbm_last_error().decode())


_initialised_device = None


def ensure_init(device: int = 0) -> None:
    """fsilbm_init once per process; raises FsilbmError if there is no usable B200."""
    global _initialised_device
    if _initialised_device is None:
        check(lib().fsilbm_init(device))
        _initialised_device = device
    elif _initialised_device != device:
        raise FsilbmError(1, f"process already bound to cuda:{_initialised_device}")
