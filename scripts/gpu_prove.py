"""Development probe (run under gpurun): Proof::prove on a synthetic SP1-shaped circuit, stage times.
SRS = random group elements (timing only; tests/test_gpu_synth.py proves with a real SRS and verifies)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import numpy as np
import dvpari, synth
ctx = dvpari.Context(0)
for lg in [int(a) for a in sys.argv[1:]] or [20]:
    t0 = time.time()
    circ = synth.synth_r1cs(lg)
    inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"], circ["coeff"], circ["coeffs_mont"])
    w = inst.synth_solve(synth.synth_assignment(circ), circ["nlevels"])
    t1 = time.time()
    dom = dvpari.Domain(ctx, lg + 1)
    t2 = time.time()
    n, k = circ["n"], circ["k"]
    ctx.srs_random(0, circ["nwires"], 1); ctx.srs_random(1, n, 2); ctx.srs_random(2, 4 * n, 3)
    prover = dvpari.Prover(ctx, dom, inst, 0, 1, 2)
    t3 = time.time()
    print(f"2^{lg}: circuit+witness {t1-t0:.1f}s domain {t2-t1:.2f}s srs {t3-t2:.2f}s  terms={sum(len(x) for x in circ['wire'])}", flush=True)
    ref = None
    for rep in range(4):
        t = time.perf_counter()
        proof = prover.prove(w[1:1 + k], w[1 + k:])
        dt = time.perf_counter() - t
        ref = ref or proof
        assert proof == ref
        st = prover.last_times()
        print(f"  prove #{rep}: {dt*1e3:.1f} ms  " + " ".join(f"{a}={b:.2f}" for a, b in st.items()), flush=True)
    prover.close(); inst.close(); dom.close()
    for s in range(3): ctx.srs_free(s)
