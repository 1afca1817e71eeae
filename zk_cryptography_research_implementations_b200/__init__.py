"""B200-native multilinear sumcheck / GKR prover hot path (drop-in for the reference's prover API).

The product is the CUDA library `libzkb200.so` (csrc/, C-ABI in include/zk_sumcheck.h); this package
is the thin host-side mirror of the reference's Rust API over it.  No CPU fallback.
"""
from .core import (BLS12_381_FR, BN254_FQ, BN254_FR, Context, DeviceTable, ReferencePanic, ZkError, fe_binop,
                   fe_from_int, fe_from_ints, fe_to_ints, synthetic_table_ints)

__all__ = ["BN254_FQ", "BN254_FR", "BLS12_381_FR", "Context", "DeviceTable", "ReferencePanic", "ZkError", "fe_binop",
           "fe_from_int", "fe_from_ints", "fe_to_ints", "synthetic_table_ints"]
