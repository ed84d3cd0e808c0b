"""Table layout: MSM time against the number of window tables W (knob msm_table_windows), persistent path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [16, 18, 19, 20, 21]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
    ctx.set("msm_table_windows", 0)
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    w0 = ctx.msm_stats()["windows"]
    row = []
    for W in range(max(9, w0 - 2), w0 + 3):
        ctx.set("msm_table_windows", W)
        ctx.multi_scalar_mul_device(d, n, 0)  # builds the tables
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
        assert out == ref
        st = ctx.msm_stats()
        row.append(f"W={W}(c={st['window_bits']}) {best*1e3:.3f}")
    print(f"tables 2^{lg} default W={w0}: " + "  ".join(row), flush=True)
    ctx.set("msm_table_windows", 0)
    ctx.dev_free(d); ctx.srs_free(0)
