"""ctypes binding of libgpb.so (include/gpb.h).  There is no CPU fallback: if the library is missing or CUDA is not
available every compute entry point raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgpb.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

# every symbol include/gpb.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gpb_version": (ctypes.c_int, []),
    "gpb_last_error": (ctypes.c_char_p, []),
    "gpb_launch_count": (ctypes.c_longlong, []),
    "gpb_program_create": (ctypes.c_int, [c_int32_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_pp]),
    "gpb_program_num_hp": (ctypes.c_int, [ctypes.c_void_p]),
    "gpb_program_destroy": (None, [ctypes.c_void_p]),
    "gpb_program_is_specialised": (ctypes.c_int, [ctypes.c_void_p]),
    "gpb_program_jit_note": (ctypes.c_char_p, [ctypes.c_void_p]),
    "gpb_jit_available": (ctypes.c_int, []),
    "gpb_jit_set_nvrtc_path": (ctypes.c_int, [ctypes.c_char_p]),
    "gpb_jit_source": (ctypes.c_int, [c_int32_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t,
                                      ctypes.POINTER(ctypes.c_size_t)]),
    "gpb_jit_cubin": (ctypes.c_int, [c_int32_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p,
                                     ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "gpb_assemble": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                    ctypes.c_void_p]),
    "gpb_plan_create": (ctypes.c_int, [ctypes.c_int, c_void_pp, c_int64_p, ctypes.c_int, c_void_pp]),
    "gpb_plan_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p]),
    "gpb_plan_bind": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "gpb_plan_buffer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_void_pp,
                                       ctypes.POINTER(ctypes.c_size_t), c_int64_p]),
    "gpb_plan_eval": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "gpb_plan_eval_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_void_pp, c_void_pp, c_void_pp,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    "gpb_plan_set_grad_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_double]),
    "gpb_plan_last_terms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "gpb_plan_input_layout": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t),
                                             ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]),
    "gpb_plan_destroy": (None, [ctypes.c_void_p]),
    "gpb_gemm": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_double, ctypes.c_double, ctypes.c_void_p]),
    "gpb_zero_upper": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "gpb_symmetrize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "gpb_dist_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "gpb_dist_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_pp]),
    "gpb_dist_loopback_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_pp]),
    "gpb_dist_destroy": (None, [ctypes.c_void_p]),
    "gpb_plan_create_dist": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, c_void_pp]),
    "gpb_plan_create_dist_columns": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, c_void_pp]),
    "gpb_dist_owner": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "gpb_dist_col_width": (ctypes.c_int, [ctypes.c_int]),
    "gpb_dist_owner_w": (ctypes.c_int, [ctypes.c_int] * 5),
    "gpb_dist_owned_cols": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p,
                                           ctypes.c_int]),
    "gpb_dist_panel_segments": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p, c_int_p, c_int_p]),
    "gpb_trtri_schedule": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, c_int_p,
                                          ctypes.c_int]),
    "gpb_debug_diag_clocks": (ctypes.c_int, [ctypes.POINTER(ctypes.c_longlong)]),
    "gpb_microbench": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "gpb_trace_begin": (ctypes.c_int, [ctypes.c_void_p]),
    "gpb_trace_end": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
}

_lib = None


class GpbError(RuntimeError):
    pass


def load():
    """Loads libgpb.so (building it is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpbError("libgpb.so not found at %s - run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().gpb_last_error().decode("utf-8", "replace")
        raise GpbError("%s failed with code %d: %s" % (what or "libgpb call", rc, msg))
