"""Run a few MSMs of 2^lg points (used as the short command profiled under ncu)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
from oracle import oracle as O

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1 << lg
G = O.generator()
ctx = dvpari.Context(0)
pts = O.chain_points(n, O.pt_mul(G, 0x1234567), O.pt_mul(G, 0x7654321))
ctx.srs_load(0, O.encode_batch(pts))
sc = dvpari.random_fr_mont(n, 5)
d = ctx.dev_alloc(n * 32)
ctx.dev_upload(d, sc)
for r in range(reps):
    t0 = time.time()
    out = ctx.multi_scalar_mul_device(d, n, 0)
    print(f"msm 2^{lg}: {1e3*(time.time()-t0):.2f} ms {out.hex()[:16]}")
