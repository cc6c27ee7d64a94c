"""ctypes binding of libni_b200.so (C ABI declared in include/ni_b200.h).

The product path has no CPU fallback: if the CUDA extension is missing the import fails, and
without a usable GPU every compute call raises `NiError` (NI_ERR_CUDA).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / 'libni_b200.so'

NI_OK, NI_ERR_INVALID, NI_ERR_CUDA, NI_ERR_NVRTC, NI_ERR_UNSUPPORTED, NI_ERR_NOMEM, NI_ERR_REC_OVERFLOW = \
    0, -1, -2, -3, -4, -5, -6
RK23, RK45 = 0, 1
MODE_ADVANCED, MODE_FAST, MODE_SCIPY = 1, 2, 4
(F_T, F_Y, F_AUX, F_H_ABS, F_STEP_SIZE, F_KLAST, F_STATUS, F_N_ACCEPTED, F_N_REJECTED, F_RUNNING, F_T_BOUND,
 F_DIRECTION, F_PARAMS, F_COND_HIT) = range(14)
F_REC_TRAJ, F_REC_STEP, F_REC_T, F_REC_Y, F_REC_AUX = 20, 21, 22, 23, 24  # device_ptr only (record log)
ST_RUNNING, ST_FINISHED, ST_FAILED, ST_NONFINITE = 0, 1, 2, 3
COND_Y_GT, COND_Y_LT, COND_AUX_GT, COND_AUX_LT = 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); mirrors include/ni_b200.h one to one (tests check the export list)
SIGNATURES = {
    'ni_version': (C.c_int, []),
    'ni_last_error': (C.c_char_p, []),
    'ni_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'ni_module_builtin': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    'ni_module_compile': (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p,
                                    C.POINTER(_vp)]),
    'ni_module_compile_stencil': (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(_vp)]),
    'ni_module_compile_component': (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(_vp)]),
    'ni_module_compile_large': (C.c_int, [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_char_p, C.POINTER(_vp)]),
    'ni_module_info': (C.c_int, [_vp] + [C.POINTER(C.c_int)] * 6),
    'ni_module_layout': (C.c_int, [_vp, C.POINTER(C.c_int)]),
    'ni_module_log': (C.c_char_p, [_vp]),
    'ni_module_destroy': (C.c_int, [_vp]),
    'ni_ensemble_create': (C.c_int, [_vp, C.c_int, C.c_int64, C.POINTER(_vp)]),
    'ni_ensemble_destroy': (C.c_int, [_vp]),
    'ni_ensemble_set_stream': (C.c_int, [_vp, _vp]),
    'ni_ensemble_upload': (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int]),
    'ni_ensemble_reset': (C.c_int, [_vp, _dp, _dp, C.c_double, C.c_double]),
    'ni_ensemble_init': (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, _dp, _dp, C.c_double, C.c_double]),
    'ni_ensemble_step': (C.c_int, [_vp, _i64p]),
    'ni_ensemble_run': (C.c_int, [_vp, _i64p]),
    'ni_ensemble_run_async': (C.c_int, [_vp]),
    'ni_ensemble_sync': (C.c_int, [_vp]),
    'ni_ensemble_poll': (C.c_int, [_vp, _i64p, _i64p]),
    'ni_ensemble_ff2t': (C.c_int, [_vp, C.c_double, _vp]),
    'ni_ensemble_ff2cond': (C.c_int, [_vp, C.c_int, C.c_int, C.c_double, _vp, _i64p]),
    'ni_ensemble_ff2cond_user': (C.c_int, [_vp, C.c_char_p, _dp, C.c_int, _vp, _i64p]),
    'ni_ensemble_get': (C.c_int, [_vp, C.c_int, _vp]),
    'ni_ensemble_set': (C.c_int, [_vp, C.c_int, _vp]),
    'ni_ensemble_device_ptr': (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    'ni_ensemble_set_max_attempts': (C.c_int, [_vp, C.c_int64]),
    'ni_ensemble_compaction': (C.c_int, [_vp, C.c_int, _i64p]),
    'ni_ensemble_set_step_limits': (C.c_int, [_vp, _vp, _vp]),
    'ni_ensemble_terminal_bytes': (C.c_int, [_vp, _i64p]),
    'ni_ensemble_pack_terminal': (C.c_int, [_vp, _vp]),
    'ni_ensemble_get_terminal': (C.c_int, [_vp, _vp]),
    'ni_ensemble_pack_terminal_peers': (C.c_int, [_vp, C.POINTER(_vp), C.c_int]),
    'ni_ensemble_terminal_sink_bytes': (C.c_int, [_vp, _i64p]),
    'ni_ensemble_stiffness': (C.c_int, [_vp, C.c_int, _vp]),
    'ni_ensemble_guard_check': (C.c_int, [_vp, _i64p, _i64p]),
    'ni_ensemble_set_order': (C.c_int, [_vp, _vp]),
    'ni_ensemble_order_by_attempts': (C.c_int, [_vp]),
    'ni_ensemble_get_order': (C.c_int, [_vp, _vp]),
    'ni_ensemble_set_terminal_sink': (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int64]),
    'ni_ensemble_record_enable': (C.c_int, [_vp, C.c_int64]),
    'ni_ensemble_record_count': (C.c_int, [_vp, _i64p]),
    'ni_ensemble_record_drain': (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, _i64p]),
    'ni_ensemble_timing': (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), _i64p]),
    'ni_fp64_peak': (C.c_int, [C.c_int, C.c_int, C.c_int, _dp]),
    'ni_devmath_check': (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_uint64, _i64p]),
}


class NiError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f'[ni_b200 error {code}] {message}')
        self.code = code


_lib = None


def lib():
    """Load the CUDA extension (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f'{LIB_PATH} is missing: build the CUDA extension with '
                              '`python -m numba_integrators_b200.build` (there is no CPU fallback)')
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != NI_OK:
        raise NiError(rc, lib().ni_last_error().decode(errors='replace'))
    return rc
