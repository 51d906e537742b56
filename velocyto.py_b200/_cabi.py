"""ctypes binding of ``libvelo_b200.so`` (C ABI declared in ``include/velo_b200.h``).

This is the only place Python touches native code.  There is no fallback: if the
library is missing, or no sm_100 device is visible when a compute entry point is
called, a :class:`VeloError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VELO_B200_LIB") or os.path.join(_PKG, "libvelo_b200.so")   # override: tuning builds only

LINEAR, SQRT, LOG10 = 0, 1, 2
RULE_FULL, RULE_PARTIAL = 0, 1
TRANSFORMS = {"linear": LINEAR, "sqrt": SQRT, "log10": LOG10, "log": LOG10}


class VeloError(RuntimeError):
    """Raised when a libvelo_b200 call returns a non-zero status."""


_i64, _int, _dbl, _ptr = C.c_int64, C.c_int, C.c_double, C.c_void_p

# name -> (restype, argtypes): every symbol include/velo_b200.h declares
PROTOTYPES = {
    "velo_abi_version": (_int, []),
    "velo_last_error": (C.c_char_p, []),
    "velo_device_info": (_int, [C.POINTER(_int), C.POINTER(_int), C.POINTER(C.c_size_t), C.POINTER(_int), C.POINTER(_int)]),
    "velo_launch_count": (C.c_uint64, []),
    "velo_release_workspace": (_int, []),
    # host drop-in tier
    "velo_colDeltaCor": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _int]),
    "velo_colDeltaCorSqrt": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _int, _dbl]),
    "velo_colDeltaCorLog10": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _int, _dbl]),
    "velo_colDeltaCorpartial": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _int]),
    "velo_colDeltaCorSqrtpartial": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _int, _dbl]),
    "velo_colDeltaCorLog10partial": (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _int, _dbl]),
    "velo_colDeltaCorpartial_compact": (_int, [_int, _ptr, _ptr, _int, _ptr, _ptr, _i64, _i64, _i64, _dbl]),
    "velo_transition_prob_partial": (_int, [_int, _ptr, _ptr, _int, _ptr, _ptr, _i64, _i64, _i64, _dbl, _dbl]),
    "velo_host_sample_neighbors_numpy": (_int, [C.c_uint32, _i64, _int, _ptr, _int, _ptr, _ptr, _ptr]),
    "velo_upload_cellmajor": (_int, [_ptr, _int, _i64, _i64, _i64, _ptr, _ptr, _ptr, _i64, _ptr]),
    "velo_transition_prob_partial_sharded": (_int, [_int, _ptr, _ptr, _i64, _ptr, _ptr, _int, _i64, _ptr, _ptr,
                                                    _i64, _i64, _i64, _i64, _i64, _dbl, _dbl]),
    # device tier
    "velo_dev_pack_cellmajor": (_int, [_ptr, _int, _i64, _i64, _ptr, _i64, _i64, _ptr]),
    "velo_dev_pack_cellmajor_split": (_int, [_ptr, _int, _i64, _i64, _ptr, _ptr, _ptr, _i64, _i64, _ptr]),
    "velo_dev_unpack_genemajor": (_int, [_ptr, _i64, _i64, _i64, _ptr, _int, _ptr]),
    "velo_dev_i64_to_i32": (_int, [_ptr, _ptr, _i64, _ptr]),
    "velo_dev_cell_stats": (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr]),
    "velo_dev_coldeltacor": (_int, [_int, _int, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64,
                                    _i64, _i64, _i64, _i64, _i64, _dbl, _ptr]),
    "velo_dev_coldeltacor_ex": (_int, [_int, _int, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64,
                                       _i64, _i64, _i64, _i64, _i64, _dbl, _ptr]),
    "velo_dev_coldeltacor_tc": (_int, [_ptr, _ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i64, _i64, _ptr, _ptr, _ptr]),
    "velo_coldeltacor_tc_workspace_bytes": (C.c_size_t, [_i64, _i64, _i64]),
    "velo_set_tensor_cores": (None, [_int]),
    "velo_get_tensor_cores": (_int, []),
    "velo_dev_scatter_dense": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _ptr]),
    "velo_dev_transition_prob": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _dbl, _ptr]),
    "velo_dev_transition_prob_ex": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _dbl, _int, _ptr]),
    "velo_dev_fit_gammas": (_int, [_int, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _dbl, _dbl,
                                   _ptr, _ptr, _ptr, _ptr, _ptr]),
    "velo_dev_fit_gammas_ex": (_int, [_int, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _dbl, _dbl, _ptr, _ptr,
                                      _ptr, _ptr, _ptr, _ptr, _ptr]),
    "velo_dev_fit_constraints": (_int, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _ptr]),
    "velo_dev_row_percentiles": (_int, [_ptr, _i64, _i64, _ptr, _int, _ptr, _ptr]),
    "velo_dev_fit_weights": (_int, [_int, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _dbl, _dbl, _ptr, _i64, _ptr]),
    "velo_dev_fit_weights_ex": (_int, [_int, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _dbl, _dbl, _dbl, _ptr, _i64, _ptr]),
    "velo_dev_velocity_chain": (_int, [_ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _i64, _int, _dbl, _dbl, _int, _int,
                                       _dbl, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "velo_dev_velocity_threshold": (_int, [_ptr, _i64, _ptr, _ptr, _i64, _i64, _dbl, _ptr, _ptr]),
    "velo_dev_delta_transform": (_int, [_ptr, _ptr, _i64, _i64, _dbl, _int, _dbl, _ptr]),
    "velo_dev_extrapolate": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _dbl, _int, _ptr]),
    "velo_dev_logratio": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _dbl, _dbl, _int, _ptr]),
    "velo_dev_row_cosine_scale": (_int, [_ptr, _ptr, _i64, _i64, _i64, _dbl, _ptr, _ptr]),
    "velo_dev_patch_corr": (_int, [_ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _int, _ptr, _ptr]),
    "velo_dev_embedding_shift": (_int, [_ptr, _i64, _ptr, _i64, _ptr, _int, _i64, _i64, _i64, _ptr, _ptr]),
    "velo_dev_cell_sums": (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr]),
    "velo_dev_size_normalize": (_int, [_ptr, _i64, _i64, _i64, _ptr, _dbl, _int, _ptr, _ptr, _ptr]),
    "velo_dev_knn_smooth": (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _int, _ptr]),
    "velo_dev_knn_smooth_csr": (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _int, _ptr]),
    "velo_dev_csr_to_cellmajor": (_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _ptr, _i64, _ptr]),
    "velo_dev_csr_cell_sums_scale": (_int, [_ptr, _ptr, _i64, _ptr, _ptr, _ptr]),
    "velo_dev_sample_neighbors": (_int, [_ptr, _i64, _int, _ptr, _int, C.c_uint64, _ptr, _ptr, _ptr]),
    "velo_dev_permute_rows_nsign": (_int, [_ptr, _ptr, _i64, _i64, _i64, C.c_uint64, _ptr]),
    "velo_dev_knn": (_int, [_ptr, _i64, _int, _int, _int, _ptr, _ptr, _ptr]),
    "velo_dev_knn_query": (_int, [_ptr, _i64, _int, _ptr, _i64, _int, _ptr, _ptr, _ptr]),
    "velo_dev_grid_flow": (_int, [_ptr, _ptr, _i64, _int, _ptr, _int, _dbl, _ptr, _ptr, _ptr]),
    "velo_dev_knn_range": (_int, [_ptr, _i64, _int, _int, _int, _i64, _i64, _ptr, _ptr, _ptr]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and attach prototypes (fails loudly when it is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VeloError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "or `make -C velocyto.py_b200/csrc` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().velo_last_error().decode("utf-8", "replace")
        raise VeloError(f"{what or 'libvelo_b200'} failed (status {status}): {msg}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)


def device_info() -> dict:
    sm, smem, hbm, maj, mnr = _int(), _int(), C.c_size_t(), _int(), _int()
    call("velo_device_info", C.byref(sm), C.byref(smem), C.byref(hbm), C.byref(maj), C.byref(mnr))
    return {"sm_count": sm.value, "smem_optin": smem.value, "hbm_bytes": hbm.value, "cc": (maj.value, mnr.value)}


def launch_count() -> int:
    return int(load().velo_launch_count())
