"""Window size sweep, plain layout (the layout of dvp_msm_adhoc and of slots below 2^15 points), persistent path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
ctx.set("msm_tables", 0)
for lg in [int(a) for a in sys.argv[1:]] or [10, 12, 13, 14, 15, 16, 17, 18]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
    ctx.set("msm_window_bits", 0)
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    c0 = ctx.msm_stats()["window_bits"]
    row = []
    for c in range(max(4, c0 - 3), min(20, c0 + 5) + 1):
        ctx.set("msm_window_bits", c)
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
        assert out == ref
        st = ctx.msm_stats()
        row.append(f"c={c}(W={st['windows']},w={st['window_bits']}) {best*1e3:.3f}")
    print(f"plain 2^{lg} default c={c0}: " + "  ".join(row), flush=True)
    ctx.set("msm_window_bits", 0)
    ctx.dev_free(d); ctx.srs_free(0)
