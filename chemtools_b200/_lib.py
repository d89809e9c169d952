"""ctypes binding of libchemtools_b200.so (C ABI in include/chemtools_b200.h).

There is no CPU fallback: if the CUDA library has not been built the import of
any compute entry point fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchemtools_b200.so")

_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)


class CtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("chemtools_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "chemtools_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl, cs = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_char_p
    sig = {
        "ct_last_error": (cs, []),
        "ct_version": (cs, []),
        "ct_mech_load": (vp, [cs, cs, C.POINTER(C.c_int)]),
        "ct_mech_load_ex": (vp, [cs, cs, cs, C.POINTER(C.c_int)]),
        "ct_mech_destroy": (None, [vp]),
        "ct_mech_save_ctm": (i32, [vp, cs]),
        "ct_mech_n_species": (i32, [vp]),
        "ct_mech_n_reactions": (i32, [vp]),
        "ct_mech_n_elements": (i32, [vp]),
        "ct_mech_species_index": (i32, [vp, cs]),
        "ct_mech_species_name": (cs, [vp, i32]),
        "ct_mech_reaction_equation": (cs, [vp, i32]),
        "ct_mech_element_name": (cs, [vp, i32]),
        "ct_mech_n_atoms": (dbl, [vp, i32, i32]),
        "ct_mech_molecular_weights": (i32, [vp, dp]),
        "ct_mech_gas_constant": (dbl, [vp]),
        "ct_mech_compiled_arrays": (i32, [vp, ip, dp, dp, dp]),
        "ct_mech_mole_to_mass": (i32, [vp, i64, dp, dp]),
        "ct_mech_flop_model": (i32, [vp, i32, dp]),
        "ct_input_parse": (vp, [cs, C.POINTER(C.c_int)]),
        "ct_input_destroy": (None, [vp]),
        "ct_input_n_case_sets": (i32, [vp, i32]),
        "ct_input_n_cases": (i32, [vp, i32, i32]),
        "ct_input_mechanism_path": (cs, [vp, i32, i32]),
        "ct_input_mechanism_description": (cs, [vp, i32, i32]),
        "ct_input_set_description": (cs, [vp, i32, i32]),
        "ct_input_case_description": (cs, [vp, i32, i32, i32]),
        "ct_input_case": (i32, [vp, i32, i32, i32, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "ct_input_case_species": (cs, [vp, i32, i32, i32, i32, dp]),
        "ct_input_case_mass_fractions": (i32, [vp, i32, i32, i32, vp, dp]),
        "ct_input_n_warnings": (i32, [vp]),
        "ct_input_warning": (cs, [vp, i32]),
        "ct_h5_create": (vp, []),
        "ct_h5_destroy": (None, [vp]),
        "ct_h5_add_dataset": (i32, [vp, cs, i32, lp, dp, cs, cs, cs]),
        "ct_h5_set_layout": (i32, [vp, i64, i32, i32]),
        "ct_h5_write": (i32, [vp, cs]),
        "ct_device_count": (i32, []),
        "ct_set_device": (i32, [i32]),
        "ct_batch_set_stream": (i32, [vp, vp]),
        "ct_batch_create": (vp, [vp, i32, i64, C.POINTER(C.c_int)]),
        "ct_batch_destroy": (None, [vp]),
        "ct_batch_n_state": (i32, [vp]),
        "ct_batch_set_ic": (i32, [vp, dp, dp, dp]),
        "ct_batch_set_ic_device": (i32, [vp, vp, vp, vp]),
        "ct_batch_set_tolerances": (i32, [vp, dbl, dbl]),
        "ct_batch_set_stop_time": (i32, [vp, dbl]),
        "ct_batch_set_max_steps": (i32, [vp, i64]),
        "ct_batch_set_max_order": (i32, [vp, i32]),
        "ct_batch_set_jacobian": (i32, [vp, i32]),
        "ct_batch_set_ignition": (i32, [vp, i32, dbl]),
        "ct_batch_set_linear_solver": (i32, [vp, i32]),
        "ct_batch_set_jacobian_clip": (i32, [vp, dbl]),
        "ct_batch_set_phase_alignment": (i32, [vp, i32]),
        "ct_batch_set_static_shape": (i32, [vp, i32]),
        "ct_mech_static_shape": (i32, [vp]),
        "ct_batch_set_matrix_pool": (i32, [vp, i32, i32]),
        "ct_batch_set_time_slicing": (i32, [vp, i32]),
        "ct_batch_set_alignment_period": (i32, [vp, i32]),
        "ct_batch_set_init_step": (i32, [vp, dbl]),
        "ct_batch_set_min_step": (i32, [vp, dbl]),
        "ct_batch_set_max_step": (i32, [vp, dbl]),
        "ct_batch_set_max_err_test_fails": (i32, [vp, i32]),
        "ct_batch_set_max_nonlin_iters": (i32, [vp, i32]),
        "ct_batch_set_max_conv_fails": (i32, [vp, i32]),
        "ct_batch_set_nonlin_conv_coef": (i32, [vp, dbl]),
        "ct_batch_reinit": (i32, [vp, dbl, vp]),
        "ct_batch_set_max_warn_messages": (i32, [vp, i32]),
        "ct_batch_set_stability_limit_detection": (i32, [vp, i32]),
        "ct_batch_set_trace": (i32, [vp, i64, i32]),
        "ct_batch_get_trace": (i32, [vp, vp, vp]),
        "ct_batch_set_timeline": (i32, [vp, i32]),
        "ct_batch_get_timeline": (i32, [vp, vp]),
        "ct_batch_set_alignment_poll_ns": (i32, [vp, i32]),
        "ct_batch_set_alignment_max_polls": (i32, [vp, i32]),
        "ct_batch_set_tp_table": (i32, [vp, i32, dp, dp, dp]),
        "ct_batch_set_v_table": (i32, [vp, i32, dp, dp, dp]),
        "ct_batch_set_order": (i32, [vp, ip]),
        "ct_batch_integrate": (i32, [vp, dbl]),
        "ct_batch_integrate_with_retry": (i32, [vp, lp, lp]),
        "ct_batch_get_stats_ex": (i32, [vp, dp]),
        "ct_batch_set_tolerances_vec": (i32, [vp, dbl, dp]),
        "ct_batch_integrate_trajectory": (i32, [vp, i32, dp]),
        "ct_batch_get_state": (i32, [vp, dp]),
        "ct_batch_get_state_device": (i32, [vp, C.POINTER(vp)]),
        "ct_batch_get_time": (i32, [vp, dp]),
        "ct_batch_get_ignition": (i32, [vp, dp]),
        "ct_batch_get_status": (i32, [vp, ip]),
        "ct_batch_get_stats": (i32, [vp, lp]),
        "ct_batch_last_kernel_ms": (dbl, [vp]),
        "ct_batch_launch_count": (i64, [vp]),
        "ct_batch_warps_per_sm": (i32, [vp]),
        "ct_host_alloc": (i32, [i64, C.POINTER(vp)]),
        "ct_host_free": (i32, [vp]),
        "ct_ensemble_create": (vp, [vp, i32, i64, i32, ip, C.POINTER(C.c_int)]),
        "ct_ensemble_destroy": (None, [vp]),
        "ct_ensemble_n_shards": (i32, [vp]),
        "ct_ensemble_set_partition": (i32, [vp, i32, i64]),
        "ct_ensemble_set_tolerances": (i32, [vp, dbl, dbl]),
        "ct_ensemble_set_stop_time": (i32, [vp, dbl]),
        "ct_ensemble_set_max_steps": (i32, [vp, i64]),
        "ct_ensemble_set_retry": (i32, [vp, i32]),
        "ct_ensemble_set_first_pass_steps": (i32, [vp, i64]),
        "ct_ensemble_set_streams": (i32, [vp, i32]),
        "ct_ensemble_set_hot_first": (i32, [vp, i32]),
        "ct_ensemble_set_configure": (i32, [vp, vp, vp]),
        "ct_ensemble_set_tp_table": (i32, [vp, i32, dp, dp, dp]),
        "ct_partition_range": (i32, [i64, i32, i32, i32, lp, lp, lp]),
        "ct_ensemble_shard_range": (i32, [vp, i32, lp, lp, lp]),
        "ct_ensemble_integrate": (i32, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "ct_ensemble_shard_info": (i32, [vp, i32, lp, dp]),
        "ct_rhs_eval": (i32, [vp, dp, dp, dp, dp, dp, dp]),
        "ct_jacobian_eval": (i32, [vp, i32, dp, dp, dp]),
        "ct_lu_solve": (i32, [i32, i64, dp, dp, dp, ip, i32]),
        "ct_rhs_time": (i32, [vp, dp, i32, dp]),
        "ct_rhs_time_cfg": (i32, [vp, dp, i32, i32, i32, dp]),
        "ct_jacobian_time": (i32, [vp, i32, dp, i32, dp]),
        "ct_lu_time": (i32, [i32, i64, i32, i32, i32, dp]),
        "ct_measure_fp64_peak": (i32, [dp]),
        "ct_roberts_selftest": (i32, [i32, dp, dbl, dp, dp, lp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    L._ct_signatures = sig
    _lib = L
    return L


def check(code):
    """Raise on negative codes (failure classes); pass 0 / 1 through."""
    if code < 0:
        raise CtError(code, lib().ct_last_error().decode())
    return code
