import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
n = 1 << 20
ctx.srs_random(0, n, 5)
sc = dvpari.random_fr_mont(n, 6)
d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, sc)
ctx.set("timing", 1)
ref = None
for minb in (1, 2, 3):
    ctx.set("pass2_minb", minb)
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
    st = ctx.msm_stats()
    ref = ref or out
    assert out == ref
    print(f"pass2_minb={minb}: {best*1e3:.2f} ms  pass2 round0 {st['ms_pass2_round0']:.3f} ms acc {st['ms_accumulate']:.2f} red {st['ms_reduce']:.2f}")
