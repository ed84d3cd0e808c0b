"""Plain layout (msm_tables = 0), persistent kernel: rounds planned ahead (msm_preplan = 1) or round by round (0), by size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
ctx.set("msm_tables", int(os.environ.get("DVP_TABLES", "0")))
for lg in [int(a) for a in sys.argv[1:]] or [12, 14, 16, 17, 18, 19, 20]:
    n = 1 << lg
    ctx.srs_random(0, n, 5)
    d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
    ref = ctx.multi_scalar_mul_device(d, n, 0)
    row = []
    for rnd in range(2):
        for pre in (1, 0):
            ctx.set("msm_preplan", pre)
            best = 1e9
            for rep in range(6):
                t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
            assert out == ref
            row.append(f"pre={pre} {best*1e3:.3f}")
    st = ctx.msm_stats()
    print(f"tables={os.environ.get('DVP_TABLES', '0')} 2^{lg} c={st['window_bits']} W={st['windows']} launches={st['launches']}: " + ", ".join(row), flush=True)
    ctx.dev_free(d); ctx.srs_free(0)
