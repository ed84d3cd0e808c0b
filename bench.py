#!/usr/bin/env python
"""Headline benchmark: sect233k1 MSM points/s at 2^20 points per GPU (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm on host cores

A step = a batch of BATCH (default 8) multi_scalar_mul calls (curve.rs:141-158), each of 2^LG uniformly random Fr
scalars (a different vector per call) against the 2^LG resident SRS points of the GPU -- the batch makes the timed region
of the driver's 20 steps longer than a second.  With N > 1 (torchrun) every rank owns a contiguous point range of
N*2^LG-point MSMs (weak scaling); the partial sums are all-gathered over NCCL and folded.
`value` times the device-resident calls (scalars already in HBM); `e2e` times the host-buffer C-ABI
call (scalars copied from pinned host memory inside the timed region, 30-byte results back).

The second half of BASELINE.json's metric -- DV-Pari prove ms at 2^22 constraints -- is measured in the same
run and reported under "prove": Proof::prove (proving.rs:426-688) through dvp_prove on a synthetic SP1-shaped
R1CS (dv-pari_b200/synth.py), witness in host memory, 118-byte proof back, with the per-stage split, the
ECFFT extend and the R1CS row evaluation rates against their rooflines.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))

METRIC = "sect233k1 MSM points/s"
UNIT = "points/s"
# Static facts about the dominant kernel, see DESIGN.md 4.2 / profiles/README.md (r2c).  At 2^20 points it is
# k_accumulate: ALL tree rounds of the bucket accumulation in one persistent launch (pass 1, per-warp inversion, pass 2
# per round; the rounds are planned ahead of the launch by k_pa_*).  Per batched affine addition it moves,
# algorithmically: descriptor 2 x 16 B read, x coordinates 2 x 32 B (pass 1), prefix product 32 B written + 32 B read,
# two points 128 B (pass 2), one point out 64 B.
ACC_BYTES_PER_ADD = 32 + 64 + 64 + 128 + 64
# dram__bytes_read.sum + dram__bytes_write.sum of one k_accumulate launch / its additions (profiles/r2s_ncu_accumulate_raw.csv:
# 5.55 GB + 1.27 GB for 13.5 M additions: the 64-byte gathers of the bucket-sorted round 0 touch whole sectors)
ACC_DRAM_BYTES_PER_ADD_NCU = 505
# ALU-pipe (LOP3 / SHF / ...) thread-instructions per addition, from ncu's own counters of the launch this line times
# (profiles/r2s_ncu_accumulate_raw.csv: sm__pipe_alu_cycles_active 66.4 % x 10.82 M active cycles / 2 x 592 schedulers =
# 2.13e9 warp-instructions of 4.30e9 executed, for 13 500 343 additions; k_pass2<16,1>: profiles/r2f_ncu_pass2_raw.csv).  The field product is now 32-bit
# multiply-adds only (gf233_mul2.cuh): ~900 LOP3 on the ALU pipe and ~890 IMAD on the FMA-heavy pipe per product, which
# dual-issue, so the ALU pipe (0.5 warp-instructions per clock and scheduler) is the pipe that binds.
ACC_ALU_INSTR_PER_ADD = 5044
PASS2_ALU_INSTR_PER_ADD = 4070
ALU_PIPE_PEAK = 1.83e13  # measured on this pool's B200: LOP3 alone runs at this many thread-instr/s
                         # (sm__pipe_alu_cycles_active = 99.9 %), profiles/r2c_ncu_pipebench2_raw.csv


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def cpu_msm_sample(ctx_or_none, lg_sample, seed, nthreads=0):
    """The oracle's restatement of the reference algorithm (per-point scalar multiplication + sum,
    curve.rs:141-158) on host cores, on a bounded sample.  Returns (points/s, cores, seconds, result)."""
    from oracle import oracle as O
    import dvpari

    n = 1 << lg_sample
    sc = dvpari.random_fr_mont(n, seed)
    if ctx_or_none is not None:
        pts, bad = O.decode_batch(ctx_or_none.srs_read(0, 0, n))
        assert bad < 0
    else:
        pts = O.mul_batch(O.generator(), dvpari.random_fr_mont(n, seed + 1))
    cores = os.cpu_count() if nthreads <= 0 else nthreads  # explicit: torchrun exports OMP_NUM_THREADS=1
    t0 = time.perf_counter()
    res = O.msm(sc, pts, cores)
    dt = time.perf_counter() - t0
    return n / dt, cores, dt, O.pt_encode(res), sc


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    import __graft_entry__ as g

    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "all"])
    # a step of the workload is a batch of MSMs of 2^lg points; the CPU arm times ONE 2^cpu_lg-point MSM per step
    # (a bounded sample of the batch: the algorithm's cost is linear in the points) and reports points/s
    lg_s = args.cpu_lg
    for _ in range(args.warmup):
        cpu_msm_sample(None, min(lg_s, 10), 11)
    tot, dt = 0, 0.0
    for s in range(args.steps):
        pps, cores, dts, _, _ = cpu_msm_sample(None, lg_s, 100 + s)  # dts: the MSM alone, not the fixture set-up
        tot += 1 << lg_s
        dt += dts
    v = tot / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (GF(2^233) via PCLMULQDQ)", "data": "synthetic",
        "config": {"workload": workload_name(args.lg, max(1, args.batch)),
                   "note": "C restatement of the reference algorithm (the Rust crate cannot be built offline); "
                           f"each step times one 2^{lg_s}-point MSM of the batch (bounded sample; cost linear in the points)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{args.steps} x 2^{lg_s} points, per-point tau-adic (width-4 TNAF) scalar mul + sum, OpenMP all cores"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ECFFT butterflies on 29-bit limbs (fr29.cuh), IMAD.WIDE per butterfly from cuobjdump of k_extend_level: decompose
# (one dot product + one multiply-add) 273, recombine (two multiply-adds) 209, the top recombine level (plain 2x2) 337.
# IMAD.WIDE issues once per 4 cycles per SMSP: 9.2e12 thread-instr/s measured (scripts/pipebench2.cu, profiles/README.md)
EXT_WIDE_DOWN, EXT_WIDE_UP, EXT_WIDE_TOP = 273, 209, 337
IMAD_WIDE_PEAK = 9.2e12


def extend_work(n, lg, polys=3):
    """(IMAD.WIDE thread-instructions, field products) of `polys` extends of n = 2^lg points."""
    half = polys * (n // 2)
    return (half * (lg * EXT_WIDE_DOWN + (lg - 1) * EXT_WIDE_UP + EXT_WIDE_TOP), half * (3 * lg + 2 * (lg - 1) + 4))


def cpu_prove_sample(ctx, args, lg=16):
    """The oracle's restatement of Proof::prove (all host cores) next to dvp_prove on the same 2^16-constraint circuit,
    SRS and witness: a bounded sample of the prove workload (the 2^22 case would take minutes on the CPU).  The two
    proofs must be the same 118 bytes."""
    import dvpari
    import synth
    from oracle import oracle as O

    circ = synth.synth_r1cs(lg, seed=0xD5A10003 + lg)
    inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"],
                               circ["coeff"], circ["coeffs_mont"])
    w = inst.synth_solve(synth.synth_assignment(circ), circ["nlevels"])
    r1cs = O.R1CS.from_arrays(circ["coeffs_mont"], circ["rowptr"], circ["wire"], circ["coeff"], circ["nrows"],
                              circ["k"], circ["nwires"])
    od = O.Domain(lg + 1)
    td = O.trapdoor(0xD5A10005, 0xD5A10006, 0xD5A10007)
    srs = O.Srs(r1cs, od, td)
    cores = os.cpu_count()
    t0 = time.perf_counter()
    want, rc, _ = O.prove(r1cs, od, srs, w, nthreads=cores)
    cpu_ms = 1e3 * (time.perf_counter() - t0)
    assert rc == 0
    dom = dvpari.Domain(ctx, lg + 1)
    for slot, pts in ((4, srs.g_m30()), (5, srs.g_q30()), (6, srs.g_k30())):
        ctx.srs_load(slot, pts)
    prover = dvpari.Prover(ctx, dom, inst, 4, 5, 6)
    k = circ["k"]
    got = prover.prove(w[1:1 + k], w[1 + k:])
    assert got == want, "device proof differs from the oracle's on the CPU-baseline sample"
    t0 = time.perf_counter()
    for _ in range(3):
        prover.prove(w[1:1 + k], w[1 + k:])
    gpu_ms = 1e3 * (time.perf_counter() - t0) / 3
    prover.close()
    inst.close()
    dom.close()
    for sl in (4, 5, 6):
        ctx.srs_free(sl)
    return {"constraints": 1 << lg, "cpu_ms_per_proof": cpu_ms, "cores": cores, "kind": "port",
            "gpu_ms_per_proof_same_input": gpu_ms, "proofs_equal": True,
            "sample": f"2^{lg}-constraint synthetic circuit, oracle dv_prove (per-point scalar multiplications, recursive extend, "
                      "serial row loop) on all host cores"}


def prove_section(ctx, args, rank, world, sync_all, timed):
    """Proof::prove at 2^prove_lg constraints: ms per proof and the stage split.  With N ranks the same proof is
    made by all of them together (strong scaling): every rank holds 1/N of g_m / g_q / g_k."""
    import numpy as np

    import dvpari
    import synth

    lg = args.prove_lg
    circ = synth.synth_r1cs(lg)
    inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"],
                               circ["coeff"], circ["coeffs_mont"])
    w = inst.synth_solve(synth.synth_assignment(circ), circ["nlevels"])
    dom = dvpari.Domain(ctx, lg + 1)
    n, k = circ["n"], circ["k"]
    # a real SRS from a fixed trapdoor, generated on the device (dvp_setup: this rank's ranges of g_m / g_q / g_k),
    # so that the timed proof can be handed to the oracle's designated verifier below
    trapdoor = [0xD5A10005, 0xD5A10006, 0xD5A10007]
    t_setup = time.perf_counter()
    dvpari.setup(inst, dom, trapdoor, 1, 2, 3)
    t_setup = time.perf_counter() - t_setup
    prover = dvpari.Prover(ctx, dom, inst, 1, 2, 3)
    import torch

    # the witness sits in pinned host memory (the upload is inside every timed dvp_prove call)
    pinned = torch.empty((circ["nwires"], 4), dtype=torch.int64).pin_memory()
    wp = pinned.numpy().view(np.uint64)
    wp[:] = w
    pub, priv = wp[1:1 + k], wp[1 + k:]
    ref = prover.prove(pub, priv)  # warm-up: sizes the scratch
    prover.prove(pub, priv)
    reps = max(3, min(args.steps, 5))
    stages = {}

    def one():
        proof = prover.prove(pub, priv)
        for a, b in prover.last_times().items():
            stages[a] = stages.get(a, 0.0) + b / reps
        return proof

    dt, proof = timed(one, reps)  # barrier + synchronize on both sides, max over ranks
    ms = 1e3 * dt / reps
    assert proof == ref
    verified = None
    if rank == 0:
        from oracle import oracle as O  # checker only: the O(1) designated verifier (srs.rs:374-428)

        verified = bool(O.verify(O.trapdoor(*trapdoor), dvpari.fr_from_mont(w[1:1 + k]), proof))
        assert verified, "the oracle's verifier rejects the benchmarked proof"
    terms = int(sum(len(x) for x in circ["wire"]))
    # the row products alone (inside a proof they share the GPU with the g_m MSM).  EVERY rank calls it: with a
    # communicator the row ranges are exchanged by an all-gather inside the call.
    rows_ms = inst.eval_time(dom, w, 5)
    if rank != 0:
        prover.close()
        inst.close()
        dom.close()
        return None
    # ECFFT extend alone: 3 polynomials of n evaluations, in place on the device
    d = ctx.dev_alloc(3 * n * 32)
    ctx.dev_upload(d, dvpari.random_fr_mont(3 * n, 5))
    dom.extend_device(d, 3)
    t0 = time.perf_counter()
    for _ in range(3):
        dom.extend_device(d, 3)
    ext_ms = 1e3 * (time.perf_counter() - t0) / 3
    ctx.dev_free(d)
    # FFTree::enter / exit of a 2^20-coefficient polynomial (BASELINE config #3), host buffers, round trip checked
    elg = min(20, lg)
    plan = dvpari.EcfftPlan(ctx, elg)
    coef = dvpari.random_fr_mont(1 << elg, 9)
    ev = plan.enter(coef)
    t0 = time.perf_counter()
    ev = plan.enter(coef)
    enter_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    back = plan.exit(ev)
    exit_ms = 1e3 * (time.perf_counter() - t0)
    assert back.tobytes() == coef.tobytes(), "exit(enter(c)) != c"
    plan.close()
    cpu_prove = cpu_prove_sample(ctx, args) if world == 1 else None
    hbm_peak, _ = peaks()
    mulmods = 3 * 4 * n * lg  # 4 n log2 n per polynomial
    ext_bytes = 3 * 2 * lg * (64 * (n // 2)) + 2 * 256 * n  # data in + out per level and polynomial, matrices once per level pair
    out = {
        "n_gpus": world, "scaling": "strong" if world > 1 else "single GPU",
        "constraints": n, "rows": circ["nrows"], "terms": terms, "wires": circ["nwires"], "public_inputs": k,
        "ms_per_proof": ms, "proofs": reps, "stage_ms": stages,
        "h2d_bytes_per_proof": int(circ["nwires"] * 32), "d2h_bytes_per_proof": 118,
        "msm_points_per_proof": circ["nwires"] + 5 * n,
        # commit_p is ONE MSM over g_m | g_q (stage "msm_gq"; "msm_gm" is then empty), then the g_k MSM: all MSM points
        # of a proof over the two MSM stages, per GPU
        "msm_points_per_s_per_gpu": (circ["nwires"] + 5 * n) / world / (1e-3 * (stages["msm_gm"] + stages["msm_gq"] + stages["msm_gk"])),
        "commit_p": "one MSM over g_m | g_q (prove_joint)",
        "ecfft_extend": {"polys": 3, "n": n, "ms": ext_ms, "mulmods_per_s": mulmods / (ext_ms * 1e-3),
                         "mulmods_note": "4 n log2 n per polynomial, the reference algorithm's count (SURVEY 8d); the "
                                         "kernels execute products_per_s (position scales carried through the tree)",
                         "products_per_s": extend_work(n, lg)[1] / (ext_ms * 1e-3),
                         "imad_wide_per_s": extend_work(n, lg)[0] / (ext_ms * 1e-3),
                         "int_frac": extend_work(n, lg)[0] / (ext_ms * 1e-3) / IMAD_WIDE_PEAK,
                         "GBps": ext_bytes / (ext_ms * 1e-3) / 1e9, "hbm_frac": ext_bytes / (ext_ms * 1e-3) / 1e9 / hbm_peak,
                         "bound": "integer (IMAD.WIDE issue), see DESIGN.md 4.3"},
        "ecfft_enter_exit": {"n": 1 << elg, "enter_ms": enter_ms, "exit_ms": exit_ms, "round_trip_exact": True,
                             "note": "host buffers (2 x 32 MiB copies inside); O(n log^2 n) built from the extend butterflies"},
        "r1cs_rows": {"ms": rows_ms, "terms_per_s": terms / (rows_ms * 1e-3),
                      "GBps": (terms * 72 + 4 * n * 32) / (rows_ms * 1e-3) / 1e9,
                      "hbm_frac": (terms * 72 + 4 * n * 32) / (rows_ms * 1e-3) / 1e9 / hbm_peak,
                      "note": "stand-alone (CUDA events); bytes = 72 per term + 128 per row of output"},
        "cpu_baseline": cpu_prove,
        "srs": "generated on the device from a fixed trapdoor (dvp_setup)", "setup_s": t_setup,
        "verified_by_oracle": verified, "data": "synthetic SP1-shaped R1CS, dv-pari_b200/synth.py",
    }
    prover.close()
    inst.close()
    dom.close()
    for sl in (1, 2, 3):
        ctx.srs_free(sl)
    return out


def workload_name(lg, batch):
    return f"sect233k1 MSM, 2^{lg} points per GPU per MSM, uniform Fr scalars, batch of {batch} MSMs per step"


def extras_section(ctx, args, rank, world, timed, n):
    """The configurations of BASELINE.json that the headline step does not cover, one short measurement each:
    plain layout (no precomputed window multiples), the table build, the ad-hoc call with encoded points from the host
    (config #2: 2^16 points), and a 2^24-point MSM split over the ranks (config #5, strong scaling)."""
    import dvpari

    out = {}
    d_sc = ctx.dev_alloc(n * 32)
    ctx.dev_upload(d_sc, dvpari.random_fr_mont(n, 0xD5A10011 + rank))
    # plain layout: one bucket set per window, no tables
    ctx.set("msm_tables", 0)
    ref = ctx.msm_sharded(d_sc, 0, on_device=True, n=n)
    dt, _ = timed(lambda: ctx.msm_sharded(d_sc, 0, on_device=True, n=n), 5)
    st = ctx.msm_stats()
    out["plain_layout"] = {"points_per_s": world * n * 5 / dt, "ms_per_msm": 1e3 * dt / 5, "window_bits": st["window_bits"],
                           "windows": st["windows"], "note": "msm_tables=0: no precomputed window multiples (memory = the SRS only)"}
    ctx.set("msm_tables", 1)
    # table build: drop and rebuild the window multiples of the slot (first MSM after a reload pays it)
    pts = ctx.srs_read(0, 0, n)
    ctx.srs_load(0, pts)
    t0 = time.perf_counter()
    again = ctx.msm_sharded(d_sc, 0, on_device=True, n=n)
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.msm_sharded(d_sc, 0, on_device=True, n=n)
    t_next = time.perf_counter() - t0
    assert again == ref, "plain and table layouts disagree"
    st = ctx.msm_stats()
    out["tables"] = {"build_ms": 1e3 * (t_first - t_next), "bytes": int(st["windows"]) * n * 64, "windows": st["windows"],
                     "note": "T[j] = 2^(off_j) P for every window j, built on the first MSM of a slot"}
    ctx.dev_free(d_sc)
    if rank == 0:
        # ad-hoc form (srs.rs:422 style): encoded points AND scalars from the host on every call, 2^16 points
        m = 1 << 16
        sc = dvpari.random_fr_mont(m, 0xD5A10001)
        enc = pts[:m]
        ctx.multi_scalar_mul_adhoc(sc, enc)
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.multi_scalar_mul_adhoc(sc, enc)
        dt = (time.perf_counter() - t0) / 5
        out["adhoc_2_16"] = {"points_per_s": m / dt, "ms_per_msm": 1e3 * dt,
                             "note": "dvp_msm_adhoc: 30-byte points + scalars from host memory, decode + MSM inside the call"}
    # 2^24 points in total, point-range-sharded over the ranks (strong scaling)
    big = (1 << args.big_lg) // world
    if args.big_lg:
        ctx.srs_random(7, big, 0xD5A10024 + rank)
        d_big = ctx.dev_alloc(big * 32)
        ctx.dev_upload(d_big, dvpari.random_fr_mont(big, 0xD5A10025 + rank))
        ctx.msm_sharded(d_big, 7, on_device=True, n=big)  # builds the tables, sizes the scratch
        dt, _ = timed(lambda: ctx.msm_sharded(d_big, 7, on_device=True, n=big), 3)
        st = ctx.msm_stats()
        out[f"msm_2_{args.big_lg}"] = {"points_total": big * world, "points_per_gpu": big, "ms_per_msm": 1e3 * dt / 3,
                                       "points_per_s": big * world * 3 / dt, "scaling": "strong" if world > 1 else "single GPU",
                                       "window_bits": st["window_bits"], "windows": st["windows"], "launches": int(st["launches"])}
        ctx.dev_free(d_big)
        ctx.srs_free(7)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lg", type=int, default=20, help="log2 of the points per GPU and MSM")
    ap.add_argument("--batch", type=int, default=8, help="MSMs (independent scalar vectors) per step")
    ap.add_argument("--cpu-lg", type=int, default=20, help="log2 of the CPU-baseline sample (2^20 points = one MSM of the workload, ~ 15-30 s of CPU-core time)")
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--prove-lg", type=int, default=22, help="log2 of the constraint count of the prove measurement (0 = skip)")
    ap.add_argument("--big-lg", type=int, default=24, help="log2 of the total points of the large strong-scaled MSM (0 = skip)")
    ap.add_argument("--no-extras", action="store_true", help="skip the plain-layout / table-build / ad-hoc / 2^24 measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = dist_env()
    import numpy as np
    import torch

    import dvpari

    if not os.path.exists(dvpari.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, batch = 1 << args.lg, max(1, args.batch)
    ctx = dvpari.Context(local)
    if use_dist:
        # the library's own NCCL communicator; torch.distributed only ships the id and runs the barriers
        ids = [dvpari.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
    if args.window_bits:
        ctx.set("msm_window_bits", args.window_bits)
    # this rank's range of the N*2^lg-point SRS and of the `batch` scalar vectors (pinned host copies + device copies)
    ctx.srs_random(0, n, 0xD5A10002 + rank)
    pinned = torch.empty((batch, n, 4), dtype=torch.int64).pin_memory()
    sc_pinned = pinned.numpy().view(np.uint64)
    d_sc = []
    for b in range(batch):
        sc_pinned[b] = dvpari.random_fr_mont(n, 0xD5A10001 + 97 * b + rank)
        d_sc.append(ctx.dev_alloc(n * 32))
        ctx.dev_upload(d_sc[b], sc_pinned[b])

    def sync_all():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
        sync_all()
        dt = time.perf_counter() - t0
        if use_dist:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, out

    # dvp_msm_sharded: local MSM over this rank's point range, all-gather of the partial sums over NCCL, identical fold
    # on every rank (with one rank it is the plain MSM)
    # A step is ONE batched call (dvp_msm_sharded_batch): the MSMs of the batch are pipelined -- the upload of the next
    # scalar vector and the enqueue of its kernels overlap the current MSM, the host folds the previous partial sums
    # meanwhile -- and one all-gather carries the whole batch.  The one-call-per-MSM loop is timed beside it.
    host_vecs = [sc_pinned[b] for b in range(batch)]
    step_dev = lambda: ctx.msm_sharded_batch(d_sc, 0, on_device=True, n=n)
    step_e2e = lambda: ctx.msm_sharded_batch(host_vecs, 0)
    step_dev_single = lambda: [ctx.msm_sharded(d_sc[b], 0, on_device=True, n=n) for b in range(batch)]
    step_e2e_single = lambda: [ctx.msm_sharded(sc_pinned[b], 0) for b in range(batch)]

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ncu_range = os.environ.get("DVP_NCU_RANGE") == "1"  # ncu --profile-from-start off: capture the timed steps only
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    dt, res_dev = timed(step_dev, args.steps)
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop() if rank == 0 else None
    launches_per_msm = int(ctx.msm_stats()["launches"])  # kernels of one MSM of the timed steps
    for _ in range(2):
        step_e2e()
    dt_e2e, res_e2e = timed(step_e2e, args.steps)
    assert res_dev == res_e2e, "device-resident and host-buffer calls disagree"
    nsingle = max(2, min(5, args.steps))
    dt_dev1, res_dev1 = timed(step_dev_single, nsingle)
    dt_e2e1, res_e2e1 = timed(step_e2e_single, nsingle)
    assert res_dev1 == res_dev and res_e2e1 == res_dev, "batched and one-call-per-MSM results disagree"

    # one instrumented MSM: stage split and the dominant kernel's duration (CUDA events on its launching stream inside
    # the library; a single lane here so that no other stream shares the GPU with the kernel being timed)
    ctx.set("timing", 1)
    ctx.set("msm_lanes", 1)
    ctx.multi_scalar_mul_device(d_sc[0], n, 0)  # sizes the single lane's scratch
    ctx.multi_scalar_mul_device(d_sc[0], n, 0)
    st = ctx.msm_stats()
    ctx.set("timing", 0)
    ctx.set("msm_lanes", 0)

    extras = None if args.no_extras else extras_section(ctx, args, rank, world, timed, n)
    prove = prove_section(ctx, args, rank, world, sync_all, timed) if args.prove_lg else None

    if rank == 0:
        hbm_peak, which = peaks()
        k_ms, k_adds = st["ms_pass2_round0"], st["adds_round0"]
        persistent = int(st["launches"]) < 40  # k_accumulate ran (the separate-launch path has > 60 launches)
        kernel = ("k_accumulate (all tree rounds of the bucket accumulation, one persistent cooperative launch)" if persistent
                  else "k_pass2<16,1> (pass 2 of round 0 of the bucket accumulation)")
        bytes_per_add = ACC_BYTES_PER_ADD if persistent else 240
        dram_per_add = ACC_DRAM_BYTES_PER_ADD_NCU if persistent else 385
        instr_per_add = ACC_ALU_INSTR_PER_ADD if persistent else PASS2_ALU_INSTR_PER_ADD
        gbps = k_adds * bytes_per_add / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        port = k_adds * instr_per_add / (k_ms * 1e-3) if k_ms > 0 else 0.0
        # bounded CPU baseline on the same SRS points (first 2^cpu_lg of rank 0's range), checked against the GPU
        lg_s = min(args.cpu_lg, args.lg)
        pps, cores, cdt, cpu_res, sc_s = cpu_msm_sample(ctx, lg_s, 0xD5A10009)
        gpu_res = ctx.multi_scalar_mul(sc_s, 0)
        assert gpu_res == cpu_res, "GPU MSM differs from the CPU oracle on the baseline sample"
        pts_per_step = world * n * batch
        config = {"workload": workload_name(args.lg, batch),
                  "window_bits": st["window_bits"], "windows": st["windows"], "precomputed_tables": bool(st["tables"]),
                  "srs": "resident in HBM (decoded once), window multiples built on first use",
                  "rounds": [st["rounds_main"], st["rounds_a"], st["rounds_b"]], "launches_per_msm": launches_per_msm,
                  "l2": "per-MSM working set (sort keys, ping-pong point buffers, prefix products: >1 GB at 2^20) exceeds the 126 MB L2; "
                        "every MSM of a step has its own scalar vector",
                  "parallelism": f"point-range sharding x{world}, NCCL all-gather of 80-byte partial sums, fold on every rank" if world > 1 else "single GPU",
                  "stage_ms": {"recode_sort": st["ms_recode_sort"], "accumulate": st["ms_accumulate"],
                               "reduce": st["ms_reduce"], "tail": st["ms_tail"]}}
        config["single_call_ms_per_msm"] = {
            "device": 1e3 * dt_dev1 / (nsingle * batch), "e2e": 1e3 * dt_e2e1 / (nsingle * batch),
            "note": "one dvp_msm_sharded call per MSM (no overlap between calls: upload, MSM, host fold in sequence)"}
        config["batched_ms_per_msm"] = {"device": 1e3 * dt / (args.steps * batch), "e2e": 1e3 * dt_e2e / (args.steps * batch)}
        if extras:
            config["other_configs"] = extras
        if prove:
            # the second half of BASELINE.json's metric, kept where the driver's parsed.config keeps it
            config["prove_ms"] = prove["ms_per_proof"]
            config["prove_constraints"] = prove["constraints"]
            config["prove_stage_ms"] = prove["stage_ms"]
            config["prove_scaling"] = prove["scaling"]
        line = {
            "metric": METRIC, "value": pts_per_step * args.steps / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 (GF(2^233) carry-less arithmetic on 8x32-bit limbs; Fr 8x32-bit Montgomery)",
            "data": "synthetic",
            "config": config,
            "e2e": {"value": pts_per_step * args.steps / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": batch * n * 32,
                    "d2h_bytes_per_step": batch * (30 + st["windows"] * st["window_bits"] * 64), "ms_per_step": 1e3 * dt_e2e / args.steps},
            "gpu_launches": launches_per_msm * batch * args.steps,
            "roofline": {"bound": "integer ALU pipe (LOP3/SHF of the carry-less field product; its IMADs dual-issue on the FMA pipe; not HBM, not tensor)",
                         "kernel": kernel, "achieved": port, "peak": ALU_PIPE_PEAK, "unit": "ALU-pipe thread-instr/s",
                         "frac": port / ALU_PIPE_PEAK, "alu_instr_per_add": instr_per_add,
                         "launch_ms": k_ms, "adds_per_launch": k_adds,
                         "peak_source": "measured on this pool's B200: LOP3 alone = 1.83e13 thread-instr/s at sm__pipe_alu_cycles_active "
                                        "99.9 % (profiles/r2c_ncu_pipebench2_raw.csv); instructions per addition from ncu's ALU-pipe "
                                        "counter of the same kernel (profiles/r2s_ncu_accumulate_raw.csv: ALU pipe 66.4 % busy)",
                         "traffic": k_adds * dram_per_add,
                         "hbm_view": {"achieved": gbps, "peak": hbm_peak, "unit": "GB/s", "frac": gbps / hbm_peak,
                                      "bytes_per_add": bytes_per_add, "peak_source": which}},
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"one MSM of the workload: 2^{lg_s} of the same SRS points, oracle k233_msm (per-point tau-adic width-4 TNAF "
                                       f"scalar mul + sum, curve.rs:141-158), {cdt:.2f} s, result equal to the GPU's"},
            "clocks": clocks,
            "device_ms_per_msm": st["ms_device"],
            "prove": prove,
        }
        print(json.dumps(line))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    for d in d_sc:
        ctx.dev_free(d)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
