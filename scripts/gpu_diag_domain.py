"""Diagnostic: device domain precomputes against the oracle, repeated, before and after other contexts lived and died."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import numpy as np

import dvpari
from oracle import oracle as O


def check(ctx, lg, tag):
    od = O.Domain(lg)
    want_w = od.vanish_derivative_on_roots_mont(0)
    want_z = od.vanish_on_other_mont(0)
    O.lib().fr_batch_inv(want_w.ctypes.data_as(C.c_void_p), C.c_size_t(want_w.shape[0]))
    O.lib().fr_batch_inv(want_z.ctypes.data_as(C.c_void_p), C.c_size_t(want_z.shape[0]))
    gd = dvpari.Domain(ctx, lg)
    z, w = gd.precomputes()
    ok_l = gd.leaves().tobytes() == od.leaves_mont().tobytes()
    bw = [i for i in range(od.n) if w[i].tobytes() != want_w[i].tobytes()]
    bz = [i for i in range(od.n) if z[i].tobytes() != want_z[i].tobytes()]
    if bw or bz or not ok_l:
        print(tag, "lg", lg, "leaves", ok_l, "bad w", bw[:8], "bad z", bz[:8])
        for i in bw[:2]:
            print("  w got", dvpari.fr_from_mont(w[i:i + 1])[0], "want", dvpari.fr_from_mont(want_w[i:i + 1])[0])
    gd.close()
    return not (bw or bz) and ok_l


bad = 0
ctx = dvpari.Context(0)
for rep in range(6):
    for lg in (2, 3, 4, 5, 8):
        bad += not check(ctx, lg, f"fresh{rep}")
ctx.close()
for rep in range(4):
    c2 = dvpari.Context(0)
    d = dvpari.Domain(c2, 11)
    d.extend(dvpari.random_fr_mont(d.n, 3))
    d.close()
    for lg in (2, 3, 4, 5, 8):
        bad += not check(c2, lg, f"after{rep}")
    c2.close()
print("mismatching domains:", bad)
