"""ctypes binding of libbatman_b200.so (C ABI in include/batman_b200.h).

There is no CPU fallback: if the shared library is missing the import fails
loudly, and every non-zero status from the library raises ``BatmanB200Error``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libbatman_b200.so')

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
vp = ctypes.c_void_p

# every symbol include/batman_b200.h declares: (restype, argtypes)
SIGNATURES = {
    'bk_version': (ctypes.c_int, []),
    'bk_last_error': (ctypes.c_char_p, []),
    'bk_device_info': (ctypes.c_int, [c_int_p, c_int_p, c_int_p]),
    'bk_profile_enable': (ctypes.c_int, [ctypes.c_int]),
    'bk_profile_read': (ctypes.c_int, [c_double_p, ctypes.POINTER(ctypes.c_long)]),
    'bk_debug_fastmath': (ctypes.c_int, [vp, ctypes.c_long, vp, vp, vp]),
    'bk_gp_create': (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'bk_gp_destroy': (ctypes.c_int, [vp]),
    'bk_gp_padded_dim': (ctypes.c_int, [ctypes.c_int]),
    'bk_gp_padded_n': (ctypes.c_int, [ctypes.c_int]),
    'bk_gp_wtiles_len': (ctypes.c_size_t, [ctypes.c_int]),
    'bk_gp_set_output': (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                        ctypes.c_double, c_double_p, vp, vp, vp, ctypes.c_double, ctypes.c_double]),
    'bk_gp_set_output_general': (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, c_int_p, c_double_p, ctypes.c_int,
                                                c_double_p, c_int_p, ctypes.c_double, vp, vp, vp, ctypes.c_double,
                                                ctypes.c_double]),
    'bk_gp_wq_bytes': (ctypes.c_size_t, [ctypes.c_int]),
    'bk_gp_set_output_int8': (ctypes.c_int, [vp, ctypes.c_int, vp, vp]),
    'bk_gp_set_var_mode': (ctypes.c_int, [vp, ctypes.c_int]),
    'bk_gp_set_overlap': (ctypes.c_int, [vp, ctypes.c_int]),
    'bk_sigma_gang_size': (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    'bk_gp_set_scaler': (ctypes.c_int, [vp, c_double_p, c_double_p]),
    'bk_pack_w': (ctypes.c_int, [vp, ctypes.c_int, vp, vp]),
    'bk_gp_predict_workspace': (ctypes.c_size_t, [vp, ctypes.c_long, ctypes.c_int]),
    'bk_gp_predict': (ctypes.c_int, [vp, vp, ctypes.c_long, vp, vp, vp, ctypes.c_size_t, vp]),
    'bk_gp_predict_status': (ctypes.c_int, [vp, ctypes.c_size_t, c_int_p, vp]),
    'bk_lml_workspace': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'bk_lml_batched': (ctypes.c_int, [vp, vp, ctypes.c_int, c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p,
                                      c_double_p, c_double_p, c_double_p, c_double_p, vp, vp, ctypes.c_size_t, vp]),
    'bk_gp_factorise': (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, c_double_p, vp, vp, vp, vp, vp, ctypes.c_int,
                                       vp, vp, vp, ctypes.c_size_t, vp]),
    'bk_lml_batched_general': (ctypes.c_int, [vp, vp, ctypes.c_int, c_int_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, c_int_p, ctypes.c_int, c_int_p, c_double_p, c_double_p,
                                              c_double_p, vp, vp, ctypes.c_size_t, vp]),
    'bk_gp_factorise_general': (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p, ctypes.c_int,
                                               c_int_p, c_double_p, c_double_p, ctypes.c_double, vp, vp, vp, vp, vp,
                                               ctypes.c_int, vp, vp, vp, ctypes.c_size_t, vp]),
    'bk_tri_inverse_pack': (ctypes.c_int, [vp, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, ctypes.c_size_t, vp]),
    'bk_sobol_nsums': (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    'bk_sobol_nslots': (ctypes.c_int, []),
    'bk_design_rows': (ctypes.c_int, [ctypes.c_uint64, ctypes.c_long, ctypes.c_long, ctypes.c_int, c_int_p,
                                      c_double_p, ctypes.c_int, vp, vp]),
    'bk_sobol_accumulate': (ctypes.c_int, [vp, vp, vp, ctypes.c_uint64, c_int_p, c_double_p, ctypes.c_long,
                                           ctypes.c_long, ctypes.c_int, c_double_p, vp, vp]),
    'bk_sobol_accumulate_keep': (ctypes.c_int, [vp, vp, vp, ctypes.c_uint64, c_int_p, c_double_p, ctypes.c_long,
                                                ctypes.c_long, ctypes.c_int, c_double_p, vp, vp, vp]),
    'bk_sobol_asymptotic': (ctypes.c_int, [vp, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int, vp, vp, vp]),
    'bk_lhs_rows': (ctypes.c_int, [ctypes.c_uint64, ctypes.c_long, ctypes.c_int, c_int_p, c_double_p, vp, vp, vp]),
    'bk_moments': (ctypes.c_int, [vp, ctypes.c_long, ctypes.c_int, vp, vp, vp]),
    'bk_sobol_modes': (ctypes.c_int, [vp, vp, vp, ctypes.c_uint64, c_int_p, c_double_p, ctypes.c_long, ctypes.c_long,
                                      ctypes.c_int, vp, vp]),
    'bk_pod_inverse': (ctypes.c_int, [vp, ctypes.c_long, ctypes.c_int, vp, vp, ctypes.c_int, vp, vp]),
    'bk_pod_sobol_accumulate': (ctypes.c_int, [vp, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp,
                                               ctypes.c_int, vp, vp, vp]),
    'bk_sobol_accumulate_y': (ctypes.c_int, [vp, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, vp, vp, vp]),
    'bk_sobol_reduce_slots': (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp]),
    'bk_sobol_jackknife': (ctypes.c_int, [vp, vp, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]),
    'bk_sobol_result_len': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    'bk_sobol_finalize': (ctypes.c_int, [vp, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]),
}


class BatmanB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


_lib = None


def load():
    """Load the shared library once; raise ImportError if it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python batman_b200/csrc/build.py` "
            "(or __graft_entry__.build()); batman_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError here = header and library out of sync
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise BatmanB200Error(load().bk_last_error().decode('utf-8', 'replace'))


def as_double_array(values):
    arr = (ctypes.c_double * len(values))(*[float(v) for v in values])
    return arr


def as_int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])
