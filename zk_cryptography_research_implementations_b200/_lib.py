"""ctypes loader of libzkb200.so (the C-ABI of include/zk_sumcheck.h).

There is no CPU fallback: if the shared library is missing this raises, and if no GPU is usable
`Context()` raises.  The library is built in-tree by `zk_cryptography_research_implementations_b200.build`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZKB200_LIB selects another build of the same library (a kernel-variant experiment, build.py ZKB200_VARIANT)
LIB_PATH = os.environ.get("ZKB200_LIB") or os.path.join(_HERE, "libzkb200.so")

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p

ZK_OK, ZK_ERR_ASSERT, ZK_ERR_CUDA, ZK_ERR_ARG = 0, -1, -2, -3
FLAG_DIRECT_S1, FLAG_SKIP_ABSORB, FLAG_NCCL_EXCHANGE, FLAG_NO_CLAIM_ABSORB, FLAG_HOST_ROUNDS, FLAG_TRUSTED_CLAIM, FLAG_HOST_EXCHANGE = 1, 2, 4, 8, 16, 32, 64

# name -> (restype, argtypes); every symbol include/zk_sumcheck.h declares
SIGNATURES = {
    "zk_version": (C.c_char_p, []),
    "zk_ctx_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int]),
    "zk_ctx_create_on_stream": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, vp]),
    "zk_ctx_destroy": (None, [vp]),
    "zk_last_error": (C.c_char_p, [vp]),
    "zk_ctx_synchronize": (C.c_int, [vp]),
    "zk_ctx_set_profiling": (C.c_int, [vp, C.c_int]),
    "zk_ctx_set_tail_log": (C.c_int, [vp, C.c_int]),
    "zk_ctx_get_tail_log": (C.c_int, [vp]),
    "zk_ctx_reset_stats": (C.c_int, [vp]),
    "zk_ctx_get_stats": (C.c_int, [vp, u64p, u64p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "zk_fe_from_u64": (C.c_int, [C.c_int, C.c_uint64, u64p]),
    "zk_fe_to_canonical": (C.c_int, [C.c_int, u64p, u64p]),
    "zk_fe_from_canonical": (C.c_int, [C.c_int, u64p, u64p]),
    "zk_fe_add": (C.c_int, [C.c_int, u64p, u64p, u64p]),
    "zk_fe_sub": (C.c_int, [C.c_int, u64p, u64p, u64p]),
    "zk_fe_mul": (C.c_int, [C.c_int, u64p, u64p, u64p]),
    "zk_interpolate_evals": (C.c_int, [C.c_int, C.c_uint32, u64p, u64p]),
    "zk_univariate_evaluate": (C.c_int, [C.c_int, u64p, C.c_uint32, u64p, u64p]),
    "zk_transcript_new": (vp, []),
    "zk_transcript_free": (None, [vp]),
    "zk_transcript_append": (None, [vp, C.c_char_p, C.c_size_t]),
    "zk_transcript_sample": (None, [vp, C.c_char_p]),
    "zk_transcript_challenge": (None, [vp, C.c_int, u64p]),
    "zk_table_upload": (C.c_int, [vp, u64p, C.c_uint64, C.POINTER(vp)]),
    "zk_table_generate": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(vp)]),
    "zk_table_regenerate": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "zk_table_wrap": (C.c_int, [vp, vp, C.c_uint64, C.POINTER(vp)]),
    "zk_table_clone": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "zk_table_download": (C.c_int, [vp, vp, u64p]),
    "zk_table_upload_into": (C.c_int, [vp, vp, vp, C.c_uint64]),
    "zk_pinned_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "zk_pinned_free": (None, [vp]),
    "zk_table_len": (C.c_uint64, [vp]),
    "zk_table_device_ptr": (vp, [vp]),
    "zk_table_free": (None, [vp, vp]),
    "zk_mle_partial_evaluate": (C.c_int, [vp, vp, C.c_uint32, u64p]),
    "zk_mle_evaluate": (C.c_int, [vp, vp, u64p, C.c_uint32, u64p]),
    "zk_mle_to_bytes": (C.c_int, [vp, vp, u8p]),
    "zk_mle_scalar_mul": (C.c_int, [vp, vp, u64p, C.POINTER(vp)]),
    "zk_mle_add": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
    "zk_mle_tensor_add": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
    "zk_mle_tensor_mul": (C.c_int, [vp, vp, vp, C.POINTER(vp)]),
    "zk_sum_halves": (C.c_int, [vp, vp, u64p]),
    "zk_sumpoly_create": (C.c_int, [vp, C.POINTER(vp), C.c_uint32, C.c_uint32, C.POINTER(vp)]),
    "zk_sumpoly_free": (None, [vp, vp]),
    "zk_sumpoly_len": (C.c_uint64, [vp]),
    "zk_sumpoly_table": (vp, [vp, C.c_uint32]),
    "zk_sumpoly_reduce": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "zk_sumcheck_round_evals": (C.c_int, [vp, vp, u64p]),
    "zk_sumcheck_fold_and_evals": (C.c_int, [vp, vp, u64p, u64p]),
    "zk_prove_product": (C.c_int, [vp, vp, u64p, vp, u64p, u64p, u64p, C.c_uint32]),
    "zk_prove_basic_device": (C.c_int, [vp, vp, u64p, u64p, u64p, u64p, C.c_uint32]),
    "zk_prove_basic": (C.c_int, [vp, u64p, C.c_uint64, u64p, u64p, u64p, u64p, C.c_uint32]),
    "zk_prove_product_host": (C.c_int, [vp, u64p, C.c_uint32, C.c_uint32, C.c_uint64, u64p, vp, u64p, u64p, u64p,
                                        C.c_uint32]),
    "zk_verify_product": (C.c_int, [C.c_int, u64p, u64p, C.c_uint32, C.c_uint32, vp, u64p, u64p, C.POINTER(C.c_int)]),
    "zk_verify_basic": (C.c_int, [vp, vp, u64p, u64p, C.c_uint32, C.POINTER(C.c_int)]),
    "zk_gkr_verify": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, u64p, u64p, u64p, u64p, C.c_uint64, C.POINTER(C.c_int)]),
    "zk_circuit_evaluate": (C.c_int, [C.c_int, vp, u64p, C.c_uint64, u64p, u64p, C.c_uint64]),
    "zk_gkr_total_rounds": (C.c_uint64, [C.c_uint32]),
    "zk_gkr_prove": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint64, u64p, u64p, u64p, u64p, u64p, u64p, u64p]),
    "zk_wide_circuit_create": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_uint32), u64p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint32), u8p, C.POINTER(vp)]),
    "zk_wide_circuit_free": (None, [vp, vp]),
    "zk_wide_circuit_total_rounds": (C.c_uint64, [vp]),
    "zk_wide_circuit_output_bits": (C.c_uint32, [vp]),
    "zk_gkr_prove_wide": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint32]),
    "zk_gkr_prove_wide_device": (C.c_int, [vp, vp, vp, u64p, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint32]),
    "zk_gkr_verify_wide": (C.c_int, [vp, vp, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint64, C.c_uint32, C.POINTER(C.c_int)]),
    "zk_gkr_verify_wide_device": (C.c_int, [vp, vp, u64p, u64p, u64p, u64p, u64p, vp, C.c_uint32, C.POINTER(C.c_int)]),
    "zk_gkr_verify_wide_succinct": (C.c_int, [vp, vp, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint32, u64p, C.POINTER(C.c_int)]),
    "zk_comm_unique_id": (C.c_int, [C.c_char_p]),
    "zk_comm_init": (C.c_int, [vp, C.c_int, C.c_int, C.c_char_p]),
    "zk_comm_attach_mailboxes": (C.c_int, [vp, C.c_char_p, C.c_int]),
    "zk_comm_unlink_mailboxes": (C.c_int, [C.c_char_p]),
    "zk_comm_peer_exchange": (C.c_int, [vp]),
    "zk_comm_destroy": (C.c_int, [vp]),
    "zk_comm_rank": (C.c_int, [vp]),
    "zk_comm_world": (C.c_int, [vp]),
    "zk_prove_product_sharded": (C.c_int, [vp, vp, u64p, vp, u64p, u64p, u64p, C.c_uint32, C.c_uint64]),
    "zk_prove_basic_sharded": (C.c_int, [vp, vp, vp, u64p, u64p, u64p, u64p, C.c_uint32, C.c_uint64]),
    "zk_gkr_prove_wide_sharded": (C.c_int, [vp, vp, vp, u64p, u64p, u64p, u64p, u64p, u64p, u64p, C.c_uint32, C.c_uint64]),
    "zk_mle_evaluate_sharded": (C.c_int, [vp, vp, u64p, C.c_uint32, u64p]),
    "zk_kzg_setup_create": (C.c_int, [vp, u64p, C.c_uint32, C.POINTER(vp)]),
    "zk_kzg_setup_from_points": (C.c_int, [vp, u64p, C.c_uint32, C.POINTER(vp)]),
    "zk_kzg_setup_free": (None, [vp, vp]),
    "zk_kzg_setup_num_vars": (C.c_uint32, [vp]),
    "zk_kzg_setup_points": (C.c_int, [vp, vp, C.c_uint32, u64p]),
    "zk_kzg_commit": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p]),
    "zk_kzg_commit_device": (C.c_int, [vp, vp, vp, u64p]),
    "zk_kzg_open": (C.c_int, [vp, vp, u64p, C.c_uint64, u64p, C.c_uint32, u64p, u64p]),
    "zk_kzg_open_device": (C.c_int, [vp, vp, vp, u64p, C.c_uint32, u64p, u64p]),
    "zk_kzg_commit_sharded": (C.c_int, [vp, vp, vp, u64p]),
    "zk_kzg_open_sharded": (C.c_int, [vp, vp, vp, u64p, C.c_uint32, u64p, u64p]),
    "zk_g1_msm": (C.c_int, [vp, u64p, u64p, C.c_uint64, u64p]),
    "zk_kzg_g2_powers_of_tau": (C.c_int, [u64p, C.c_uint32, u64p]),
    "zk_kzg_verify": (C.c_int, [u64p, C.c_uint32, u64p, u64p, C.c_uint32, u64p, u64p, C.c_uint32, C.POINTER(C.c_int)]),
    "zk_pairing_product_is_one": (C.c_int, [u64p, u64p, C.c_uint32, C.POINTER(C.c_int)]),
    "zk_g1_is_on_curve": (C.c_int, [u64p]),
    "zk_g1_generator": (None, [u64p]),
    "zk_g2_generator": (None, [u64p]),
    "zk_g1_arith_probe": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "zk_arith_probe": (C.c_int, [vp, C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libzkb200.so is not built (%s). Run `python -m zk_cryptography_research_implementations_b200.build`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
