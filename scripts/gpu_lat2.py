"""Development probe: single-warp latencies (inversions, multiplication) and the per-launch cost of small tree rounds."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for mode, name in ((0, "gf_inv (squarings)"), (1, "gf_inv_tab"), (2, "gf_mul"), (3, "gf_mul_warp"), (4, "gf_inv_warp")):
    print(name, "us/op single warp:", ctx.latency_probe(mode, 200 if mode in (2, 3) else 20))
for fused in (1,):
    for lg in [int(a) for a in sys.argv[1:]] or [10, 12, 14]:
        n = 1 << lg
        ctx.srs_random(0, n, 5)
        sc = dvpari.random_fr_mont(n, 6)
        d = ctx.dev_alloc(n * 32); ctx.dev_upload(d, sc)
        ctx.set("msm_profile", 1); ctx.set("timing", 1)
        for rep in range(3):
            out = ctx.multi_scalar_mul_device(d, n, 0)
        pr = ctx.msm_profile(); st = ctx.msm_stats()
        ctx.set("msm_profile", 0); ctx.set("timing", 0)
        print(f"n=2^{lg} c={st['window_bits']} W={st['windows']} tables={st['tables']} rounds={st['rounds_main']},{st['rounds_a']},{st['rounds_b']}: " +
              "  ".join(f"{k} {1e3*v[0]/max(1,v[1]):.1f}us x{v[1]}" for k, v in pr.items()))
        tot = sum(v[0] for v in pr.values())
        best = 1e9
        for rep in range(5):
            t0 = time.perf_counter(); out2 = ctx.multi_scalar_mul_device(d, n, 0); best = min(best, time.perf_counter() - t0)
        assert out2 == out
        print(f"   {best*1e3:.3f} ms wall ({n/best:.3e} pts/s), {ctx.msm_stats()['launches']} launches; profiled single-lane sum {tot:.3f} ms")
        print(f"   stages: sort {st['ms_recode_sort']:.3f} accumulate {st['ms_accumulate']:.3f} reduce {st['ms_reduce']:.3f} tail {st['ms_tail']:.3f}")
        ctx.dev_free(d)
