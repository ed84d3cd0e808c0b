"""The short command profiled under ncu (profiles/README.md): one MSM of 2^20 points and one prove of 2^LG
constraints inside a cudaProfilerStart/Stop range; set-up (SRS, window tables, circuit, warm-up) stays outside.
    python scripts/prof_cmd.py [lg_prove] [what]      what: msm | prove | both"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import numpy as np
import torch
import dvpari, synth

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
what = sys.argv[2] if len(sys.argv) > 2 else "both"
ctx = dvpari.Context(0)
for kv in filter(None, os.environ.get("DVP_KNOBS", "").split(",")):  # e.g. DVP_KNOBS=use_accumulate=0,msm_lanes=1
    ctx.set(kv.split("=")[0], int(kv.split("=")[1]))
n = 1 << 20
if what in ("msm", "both"):
    ctx.srs_random(0, n, 5)
    d = ctx.dev_alloc(n * 32)
    ctx.dev_upload(d, dvpari.random_fr_mont(n, 6))
    for _ in range(2):
        ref = ctx.multi_scalar_mul_device(d, n, 0)
if what in ("prove", "both"):
    circ = synth.synth_r1cs(lg)
    inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"], circ["coeff"], circ["coeffs_mont"])
    w = inst.synth_solve(synth.synth_assignment(circ), circ["nlevels"])
    dom = dvpari.Domain(ctx, lg + 1)
    N, k = circ["n"], circ["k"]
    ctx.srs_random(1, circ["nwires"], 1); ctx.srs_random(2, N, 2); ctx.srs_random(3, 4 * N, 3)
    prover = dvpari.Prover(ctx, dom, inst, 1, 2, 3)
    for _ in range(2):
        pref = prover.prove(w[1:1 + k], w[1 + k:])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if what in ("msm", "both"):
    t0 = time.perf_counter(); out = ctx.multi_scalar_mul_device(d, n, 0); t1 = time.perf_counter()
    assert out == ref
    st = ctx.msm_stats()
    print(f"msm 2^20: {1e3*(t1-t0):.2f} ms, {st['launches']} launches, c={st['window_bits']} W={st['windows']} tables={st['tables']}")
if what in ("prove", "both"):
    t0 = time.perf_counter(); proof = prover.prove(w[1:1 + k], w[1 + k:]); t1 = time.perf_counter()
    assert proof == pref
    print(f"prove 2^{lg}: {1e3*(t1-t0):.2f} ms {prover.last_times()}")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
