"""Time Domain.from_fftree_file on a tree2n file of the 2^22-constraint size (2^23 leaves, FLeaves = 487 MB)."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import numpy as np

import artifacts
import dvpari

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 23
ctx = dvpari.Context(0)
t0 = time.time()
dom = dvpari.Domain(ctx, lg)
t1 = time.time()
leaves = dom.leaves()
dom.close()
f = np.zeros((2 << lg, 4), dtype=np.uint64)  # heap array: only the leaf layer is filled in
f[1 << lg:] = leaves
ident = np.zeros((0, 4), dtype=np.uint64)
path = os.path.join(tempfile.gettempdir(), "tree2n")
artifacts.write_fftree_to_file(path, dict(f=f, recombine=ident, decompose=ident))
t2 = time.time()
dom = dvpari.Domain.from_fftree_file(ctx, path)
t3 = time.time()
ev = dvpari.random_fr_mont(1 << 10, 1)
print(f"lg={lg} file={os.path.getsize(path) / 1e6:.0f} MB  domain_create {t1 - t0:.2f} s  "
      f"from_fftree_file {t3 - t2:.2f} s  n2={dom.n2}")
dom.close()
os.remove(path)
ctx.close()
