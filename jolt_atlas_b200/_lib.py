"""ctypes loader for the in-tree C-ABI library.  Fails loudly when the library is missing: there is
no Python/CPU fallback for any kernel."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libjolt_atlas_b200.so")

u64p = C.c_void_p     # Fr / Fq limb arrays travel as plain addresses (ndarray.ctypes.data): POINTER(c_uint64) conversion costs ~1.3 us per argument
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes).  Must list every symbol include/jolt_atlas_b200.h declares
# (tests/test_abi.py checks the header against this table and against the built .so).
SIGNATURES = {
    "ja_init": (C.c_int32, [C.c_int32, vpp]),
    "ja_shutdown": (None, [vp]),
    "ja_last_error": (None, [C.c_char_p, C.c_size_t]),
    "ja_sync": (C.c_int32, [vp]),
    "ja_launch_count": (C.c_uint64, [vp]),
    "ja_poly_from_fr": (C.c_int32, [vp, u64p, C.c_size_t, vpp]),
    "ja_poly_from_i32": (C.c_int32, [vp, i32p, C.c_size_t, vpp]),
    "ja_poly_from_i32_many": (C.c_int32, [vp, i32p, C.c_size_t, C.c_size_t, vpp]),
    "ja_poly_from_lookup": (C.c_int32, [vp, u64p, C.c_size_t, u32p, C.c_size_t, vpp]),
    "ja_poly_alloc": (C.c_int32, [vp, C.c_size_t, vpp]),
    "ja_poly_clone": (C.c_int32, [vp, vp, vpp]),
    "ja_poly_len": (C.c_size_t, [vp]),
    "ja_poly_to_host": (C.c_int32, [vp, vp, u64p, C.c_size_t]),
    "ja_poly_free": (None, [vp, vp]),
    "ja_poly_free_many": (None, [vp, vp, C.c_size_t]),
    "ja_bind": (C.c_int32, [vp, vp, u64p, C.c_int32]),
    "ja_bind_many": (C.c_int32, [vp, vpp, C.c_size_t, u64p, C.c_int32]),
    "ja_final_claim": (C.c_int32, [vp, vp, u64p]),
    "ja_poly_evaluate": (C.c_int32, [vp, vp, u64p, C.c_size_t, u64p]),
    "ja_eq_evals": (C.c_int32, [vp, u64p, C.c_size_t, u64p, vpp]),
    "ja_spliteq_new": (C.c_int32, [vp, u64p, C.c_size_t, C.c_int32, u64p, vpp]),
    "ja_spliteq_bind": (C.c_int32, [vp, vp, u64p]),
    "ja_spliteq_current_scalar": (C.c_int32, [vp, u64p]),
    "ja_spliteq_current_w": (C.c_int32, [vp, u64p]),
    "ja_spliteq_merge": (C.c_int32, [vp, vp, vpp]),
    "ja_spliteq_free": (None, [vp, vp]),
    "ja_round_eval": (C.c_int32, [vp, C.c_int32, vpp, C.c_size_t, vp, u64p, C.c_size_t, C.c_uint32, u64p, C.c_size_t]),
    "ja_sumcheck_prove": (C.c_int32, [vp, C.c_int32, vpp, C.c_size_t, u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint32, u64p,
                                      C.c_char_p, u32p, C.c_size_t, u64p, u32p, u64p, u64p]),
    "ja_batched_sumcheck_prove": (C.c_int32, [vp, vp, C.c_size_t, C.c_char_p, u32p, C.c_size_t, u64p, u32p, u64p]),
    "ja_eval_reduction_h": (C.c_int32, [vp, vp, u64p, C.c_size_t, C.c_size_t, u64p, C.POINTER(C.c_size_t)]),
    "ja_tensor_fold_i32": (C.c_int32, [vp, i32p, C.c_size_t, C.c_size_t, vp, C.c_int32, vpp]),
    "ja_tensor_i32_upload": (C.c_int32, [vp, i32p, C.c_size_t, C.c_size_t, vpp]),
    "ja_tensor_i32_free": (None, [vp, vp]),
    "ja_tensor_fold_resident": (C.c_int32, [vp, vp, vp, C.c_int32, vpp]),
    "ja_srs_upload": (C.c_int32, [vp, u64p, C.c_size_t, vpp]),
    "ja_srs_generate": (C.c_int32, [vp, u64p, u64p, C.c_size_t, vpp]),
    "ja_srs_precompute": (C.c_int32, [vp, vp]),
    "ja_srs_to_host": (C.c_int32, [vp, vp, C.c_size_t, C.c_size_t, u64p]),
    "ja_srs_len": (C.c_size_t, [vp]),
    "ja_srs_free": (None, [vp, vp]),
    "ja_msm_fr": (C.c_int32, [vp, vp, vp, u64p, i32p]),
    "ja_msm_fr_batch": (C.c_int32, [vp, vp, vpp, C.c_size_t, u64p, i32p]),
    "ja_msm_host": (C.c_int32, [vp, vp, C.c_size_t, vp, C.c_int32, C.c_size_t, u64p, i32p]),
    "ja_g1_sum_indexed": (C.c_int32, [vp, vp, u64p, C.c_size_t, u64p, i32p]),
    "ja_g1_sum_indexed_batch": (C.c_int32, [vp, vp, u64p, u64p, C.c_size_t, u64p, i32p]),
    "ja_onehot_upload": (C.c_int32, [vp, u64p, u64p, C.c_size_t, vpp]),
    "ja_onehot_commit": (C.c_int32, [vp, vp, vp, u64p, i32p]),
    "ja_onehot_free": (None, [vp, vp]),
    "ja_addr_upload": (C.c_int32, [vp, u32p, C.c_size_t, C.c_size_t, C.c_size_t, vpp]),
    "ja_addr_upload_many": (C.c_int32, [vp, vp, vp, vp, vp, C.c_size_t, vp]),
    "ja_addr_free": (None, [vp, vp]),
    "ja_addr_len": (C.c_size_t, [vp]),
    "ja_addr_count": (C.c_size_t, [vp]),
    "ja_addr_commit": (C.c_int32, [vp, vp, vp, u64p, i32p]),
    "ja_addr_commit_many": (C.c_int32, [vp, vp, vpp, C.c_size_t, u64p, i32p]),
    "ja_addr_gather": (C.c_int32, [vp, vp, u64p, vpp]),
    "ja_addr_ra_evals": (C.c_int32, [vp, vp, u64p, C.c_size_t, u64p]),
    "ja_addr_ra_evals_many": (C.c_int32, [vp, vp, vp, vp, C.c_size_t, vp]),
    "ja_poly_zeros": (C.c_int32, [vp, C.c_size_t, vpp]),
    "ja_rlc_add_onehot": (C.c_int32, [vp, vp, vp, u64p]),
    "ja_rlc_add_dense": (C.c_int32, [vp, vp, vp, u64p]),
    "ja_hyperkzg_open_begin": (C.c_int32, [vp, vp, vp, u64p, C.c_size_t, vpp, u64p, i32p]),
    "ja_hyperkzg_open_evals": (C.c_int32, [vp, vp, u64p, u64p]),
    "ja_hyperkzg_open_witness": (C.c_int32, [vp, vp, u64p, u64p, u64p, i32p]),
    "ja_hyperkzg_open_free": (None, [vp, vp]),
    "ja_hyperkzg_open": (C.c_int32, [vp, vp, vp, u64p, C.c_size_t, C.c_char_p, u32p, u64p, i32p, u64p, i32p, u64p]),
    "ja_set_sumcheck_shard": (C.c_int32, [vp, C.c_uint32, C.c_uint32, vp, vp]),
    "ja_set_msm_shard": (C.c_int32, [vp, C.c_uint32, C.c_uint32]),
    "ja_msm_fr_range": (C.c_int32, [vp, vp, vp, C.c_size_t, C.c_size_t, u64p, i32p]),
    "ja_round_eval_slice": (C.c_int32, [vp, C.c_int32, vpp, C.c_size_t, vp, C.c_uint32, C.c_size_t, u64p, C.c_size_t]),
    "ja_psshout_new": (C.c_int32, [vp, u64p, C.c_size_t, u64p, C.c_size_t, C.c_uint32, C.c_uint32, vpp]),
    "ja_psshout_init_phase": (C.c_int32, [vp, vp, C.c_uint32, u64p, u32p, C.c_size_t, C.c_uint32, u64p]),
    "ja_psshout_materialize_ra": (C.c_int32, [vp, vp, u64p, u64p, vpp]),
    "ja_psshout_prove_address": (C.c_int32, [vp, vp, C.c_uint32, u64p, u64p, C.c_char_p, C.POINTER(C.c_uint32), u64p, u32p, u64p, u64p, u64p, u64p, u64p]),
    "ja_psshout_tables": (C.c_int32, [vp, vp, u64p]),
    "ja_set_cache_openings": (C.c_int32, [vp, C.c_int32]),
    "ja_transcript_append_scalar_each": (None, [C.c_char_p, C.POINTER(C.c_uint32), u64p, C.c_size_t]),
    "ja_psshout_prove_identity_rc": (C.c_int32, [vp, vp, u64p, C.c_char_p, C.POINTER(C.c_uint32), u64p, u32p, u64p, u64p, u64p, u64p]),
    "ja_psshout_from_witness_rem": (C.c_int32, [vp, vp, u64p, C.c_size_t, C.c_uint32, vpp]),
    "ja_psshout_free": (None, [vp, vp]),
    "ja_witness_fused": (C.c_int32, [vp, C.c_int32, vp, vp, C.c_uint32, C.c_size_t, vpp]),
    "ja_witness_clamp_addr": (C.c_void_p, [vp]),
    "ja_witness_rem_addr": (C.c_void_p, [vp]),
    "ja_witness_to_host": (C.c_int32, [vp, vp, vp, vp, vp, vp]),
    "ja_psshout_from_witness": (C.c_int32, [vp, vp, u64p, C.c_size_t, C.c_uint32, C.c_uint32, vpp]),
    "ja_psshout_new_dev": (C.c_int32, [vp, vp, C.c_size_t, u64p, C.c_size_t, C.c_uint32, C.c_uint32, vpp]),
    "ja_witness_free": (None, [vp, vp]),
    "ja_comm_unique_id": (C.c_int32, [C.c_char_p]),
    "ja_comm_init": (C.c_int32, [vp, C.c_uint32, C.c_uint32, C.c_char_p]),
    "ja_comm_free": (None, [vp]),
    "ja_comm_allgather": (C.c_int32, [vp, vp, C.c_size_t, vp]),
    "ja_g1_sum_affine": (C.c_int32, [u64p, i32p, C.c_size_t, u64p, i32p]),
    "ja_fr_sum": (C.c_int32, [u64p, C.c_size_t, C.c_size_t, u64p]),
    "ja_transcript_new": (None, [C.c_char_p, C.c_char_p, u32p]),
    "ja_transcript_append_points": (None, [C.c_char_p, u32p, u64p, i32p, C.c_size_t]),
    "ja_transcript_append_scalars": (None, [C.c_char_p, u32p, u64p, C.c_size_t]),
    "ja_transcript_challenge_scalar": (None, [C.c_char_p, u32p, u64p]),
    "ja_transcript_challenge_scalar_powers": (None, [C.c_char_p, u32p, C.c_size_t, u64p]),
    "ja_transcript_challenge_optimized": (None, [C.c_char_p, u32p, C.c_size_t, u64p]),
    "ja_expanding_table": (C.c_int32, [u64p, C.c_size_t, C.c_int32, u64p]),
    "ja_profile_begin": (C.c_int32, [vp]),
    "ja_profile_end": (C.c_int32, [vp, u64p, C.POINTER(C.c_double), C.c_size_t]),
    "ja_profile_class_count": (C.c_int32, []),
    "ja_profile_class_name": (C.c_char_p, [C.c_int32]),
    "ja_timer_begin": (C.c_int32, [vp]),
    "ja_timer_end": (C.c_int32, [vp, C.POINTER(C.c_float)]),
    "ja_bench_kernel": (C.c_int32, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "ja_bench_fused": (C.c_int32, [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "ja_poly_random": (C.c_int32, [vp, C.c_size_t, C.c_uint32, vpp]),
    "ja_calibrate_fr_mul": (C.c_int32, [vp, C.c_int32, C.POINTER(C.c_double)]),
    "ja_test_field_ops": (C.c_int32, [vp, C.c_int32, u64p, u64p, C.c_size_t, u64p]),
}



class ScInstance(C.Structure):
    """ja_sc_instance (include/jolt_atlas_b200.h)."""
    _fields_ = [("kind", C.c_int32), ("aux_u32", C.c_uint32), ("n_polys", C.c_size_t), ("polys", C.c_void_p),
                ("host_tables", C.c_void_p), ("table_len", C.c_size_t), ("addr", C.c_void_p), ("eq_w", C.c_void_p),
                ("eq_m", C.c_size_t), ("aux_fr", C.c_void_p), ("n_aux", C.c_size_t), ("claim", C.c_uint64 * 4),
                ("out_final_claims", C.c_void_p)]


_lib = None


class JoltAtlasError(RuntimeError):
    """Raised for any non-zero status from the C ABI (the reference prover panics in the same places)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"jolt_atlas_b200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m jolt_atlas_b200.build` "
            "(this package has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        buf = C.create_string_buffer(1024)
        load().ja_last_error(buf, 1024)
        raise JoltAtlasError(status, buf.value.decode("utf-8", "replace"))
