import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for mode, nm, it in [(2, "gf_mul", 200), (0, "gf_inv (232 squarings)", 20), (1, "gf_inv_tab", 20)]:
    print(f"{nm}: {ctx.latency_probe(mode, it):.2f} us per op (single warp, dependent)")
