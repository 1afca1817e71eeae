#!/usr/bin/env python3
"""Register-resident product probes (zk_arith_probe): chained mul_acc, carry-out columns, flag-free radix 2^29."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_cryptography_research_implementations_b200 as zk  # noqa: E402

ctx = zk.Context(zk.BN254_FQ, 0)
out = {}
for kind, name in ((0, "mont_mul"), (1, "fold_by_scalar"), (2, "mul_acc chained (4 chains, one limb varies)"), (7, "mul_acc_cols carry-out slots (2 chains)"),
                   (9, "mul_acc chained, fully varying operands (2 chains)"), (8, "radix-2^29 flag-free products, fully varying operands (2 chains)"),
                   (4, "IMAD.WIDE.U32 plain"), (6, "IMAD.WIDE.U32.X chained")):
    for bps in (1, 2):
        ops, ms = C.c_double(), C.c_double()
        ctx.check(ctx.lib.zk_arith_probe(ctx.h, kind, 1500, bps, C.byref(ops), C.byref(ms)))
        out["%s @%d blocks/SM" % (name, bps)] = round(ops.value / 1e9, 1)
print(json.dumps(out, indent=1))
