"""Static SASS instruction mix of the field / MSM kernels (cuobjdump, no GPU): the per-addition instruction counts
bench.py's int_issue block is built from.  ALU pipe: LOP3, SHF, PRMT, IADD3, ISETP, SEL, ...; FMA pipe: IMAD*.
IMAD.WIDE occupies the FMA pipe for 4 cycles and the ALU pipe for 2 (profiles/README.md, pipe probes)."""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "dv-pari_b200", "csrc", "msm.o")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
def classify(c):
    wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
    fma = sum(v for k, v in c.items() if k.startswith("IMAD")) - wide
    alu = sum(v for k, v in c.items() if re.match(r"(LOP3|SHF|PRMT|IADD3|ISETP|SEL|VIADD|LEA|MOV|VIMNMX|PLOP3|ICMP)", k))
    mem = sum(v for k, v in c.items() if re.match(r"(LDG|STG|LDS|STS|LDL|STL|LDC)", k))
    return dict(imad_wide=wide, imad_other=fma, alu=alu, mem=mem, total=sum(c.values()))
out = {}
for name, c in funcs.items():
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    if any(t in short for t in ("k_pass2<16, 2>", "k_pass1<16>", "gf_mul_call", "k_binv_up", "k_binv_direct")):
        out[short] = classify(c)
print(json.dumps(out, indent=1))
p2, mul, p1 = out.get("void dvp::k_pass2<16, 2>"), out.get("dvp::gf_mul_call"), out.get("void dvp::k_pass1<16>")
if p2 and mul:
    # per addition: the loop body once + 4 calls of the out-of-line multiplier (dinv, inv update, lambda, y3)
    per_add = {k: p2[k] + 4 * mul[k] for k in ("imad_wide", "imad_other", "alu", "total")}
    per_add["alu_pipe_cycles_per_warp"] = 2 * (per_add["alu"] + per_add["imad_wide"])
    print("k_pass2 per addition (upper bound: whole kernel body counted once):", json.dumps(per_add))
