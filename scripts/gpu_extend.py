"""Development probe (run under gpurun): ECFFT extend of 3 polynomials, ms and IMAD.WIDE issue fraction."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:] if a.isdigit()] or [20, 22]:
    n = 1 << lg
    dom = dvpari.Domain(ctx, lg + 1)
    d = ctx.dev_alloc(3 * n * 32); ctx.dev_upload(d, dvpari.random_fr_mont(3 * n, 5))
    dom.extend_device(d, 3)
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter(); dom.extend_device(d, 3); best = min(best, time.perf_counter() - t0)
    wides = 3 * (n // 2) * (lg * 273 + (lg - 1) * 209 + 337)  # bench.py extend_work
    print(f"extend 3 x 2^{lg}: {best*1e3:.3f} ms  {wides/best:.3e} IMAD.WIDE/s = {wides/best/9.2e12:.3f} of peak")
    ctx.dev_free(d); dom.close()
# enter / exit (BASELINE config #3: 2^20-coefficient polynomials)
import numpy as np
for lg in (() if "--no-enter" in sys.argv else (16, 20)):
    n = 1 << lg
    t0 = time.perf_counter(); plan = dvpari.EcfftPlan(ctx, lg); t_plan = time.perf_counter() - t0
    c = dvpari.random_fr_mont(n, 9)
    ev = plan.enter(c)
    t0 = time.perf_counter(); ev = plan.enter(c); t_enter = time.perf_counter() - t0
    t0 = time.perf_counter(); back = plan.exit(ev); t_exit = time.perf_counter() - t0
    assert back.tobytes() == c.tobytes()
    print(f"2^{lg}: plan {t_plan:.2f} s, enter {t_enter*1e3:.1f} ms, exit {t_exit*1e3:.1f} ms (host buffers), exit(enter(c)) == c")
    plan.close()
