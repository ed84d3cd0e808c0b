"""Regenerates tests/golden/golden_v1.json from the oracle (run from the repo root).

The reference (Rust + un-vendored crates) cannot be run in this image, so these vectors are NOT
upstream outputs: they freeze the oracle after it was pinned against OpenSSL / brute force / the
Python blake3 package, and they record the *candidate* xsk233 encodings that a machine with cargo
should compare against `CurvePoint::to_bytes`."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O

P = O.P
rnd = random.Random(20261018)
G = O.generator()
out = {"note": "oracle-generated; xsk233 encodings are candidates (parity unpinned)", "fr": {}, "gf": [], "points": [], "ecfft": {}, "toy": {}}
out["fr"] = {"p": hex(P), "R_mod_p": hex((1 << 256) % P), "R2_mod_p": hex((1 << 512) % P)}
for _ in range(4):
    a, b = rnd.getrandbits(233), rnd.getrandbits(233)
    out["gf"].append({"a": hex(a), "b": hex(b), "mul": hex(O.gf_mul(a, b)), "sqr_a": hex(O.gf_sqr(a)), "inv_a": hex(O.gf_inv(a))})
for k in [1, 2, 3, 0xDEADBEEF, P - 1, rnd.randrange(P)]:
    pt = O.pt_mul(G, k)
    x, y = O.pt_xy(pt)
    out["points"].append({"k": hex(k), "x": hex(x), "y": hex(y), "xsk233_candidate": O.pt_encode(pt).hex()})
out["points"].append({"k": "0x0", "x": None, "y": None, "xsk233_candidate": O.pt_encode(O.pt()).hex()})
dom = O.Domain(4)
lv = dom.leaves()
ev = [rnd.randrange(P) for _ in range(8)]
out["ecfft"] = {"log_n2": 4, "leaves": [hex(v) for v in lv], "evals_on_D": [hex(v) for v in ev],
                "extend_to_Dprime": [hex(v) for v in dom.extend(ev)],
                "Z_D_at_12345": hex(dom.vanish_at(0, 12345))}
r1cs, pub, priv = O.toy_r1cs()
td_vals = [rnd.randrange(1, P) for _ in range(3)]
td = O.trapdoor(*td_vals)
srs = O.Srs(r1cs, dom, td)
proof, rc, _ = O.prove(r1cs, dom, srs, O.mont_array([1] + pub + priv))
assert rc == 0 and O.verify(td, pub, proof)
out["toy"] = {"trapdoor": [hex(v) for v in td_vals], "public": pub, "private": priv, "proof118": proof.hex(),
              "alpha": hex(O.transcript_alpha(proof[:30], pub)),
              "g_m": O.Srs.g_m30(srs).tobytes().hex(), "g_q": srs.g_q30().tobytes().hex(), "g_k": srs.g_k30().tobytes().hex()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden_v1.json"), "w"), indent=1)
print("written")
