"""Development probe (run under gpurun): MSM time vs table window width."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [20]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    sc = dvpari.random_fr_mont(n, 6)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, sc)
    ctx.set("msm_tables", 0)
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    ctx.set("msm_tables", 1)
    for cb in (0, 12, 13, 14, 15, 16, 17, 18, 20, 22):
        ctx.set("msm_table_windows", cb)
        t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); tb = time.perf_counter() - t0
        assert out == ref, cb
        for lanes in (1, 2, 3):
            ctx.set("msm_lanes", lanes)
            best = 1e9
            for rep in range(5):
                t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
            assert out == ref
            st = ctx.msm_stats()
            print(f"n=2^{lg} table_windows={cb} c={st['window_bits']} W={st['windows']} lanes={lanes}: {best*1e3:.2f} ms {n/best:.3e} pts/s rounds={st['rounds_main']},{st['rounds_a']},{st['rounds_b']} (first call incl. build {tb*1e3:.0f} ms)", flush=True)
    ctx.set("msm_lanes", 0); ctx.set("msm_table_windows", 0)
    ctx.dev_free(d)
