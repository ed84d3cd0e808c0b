//! Golden capture: runs the UNMODIFIED reference (alpenlabs/dv-pari + xs233-sys =0.2.0) and writes the byte vectors
//! that pin the one part of the port no offline check can reach -- the 30-byte xsk233 point codec and the identity of
//! `xsk233_generator` -- plus a whole toy proof and the reference's own tree file.
//!
//!   cargo run --release --manifest-path rust-shim/capture-golden/Cargo.toml -- tests/golden
//!
//! Output (layout consumed by tests/test_golden.py::test_reference_capture_*):
//!   reference_v1.json
//!     generator      hex(CurvePoint::generator().to_bytes())                         src/curve.rs:84-99
//!     neutral        hex(xsk233_neutral encoded)                                     src/io_utils.rs:253-267
//!     mulgen[]       { k: hex canonical integer, enc: hex(point_scalar_mul_gen(k).to_bytes()) }   src/curve.rs:129-137
//!     mul[]          { k, base: enc of the base point, enc: hex(point_scalar_mul(k, base)) }      src/curve.rs:113-126
//!     msm16          { scalars[16], points[16] (enc), result }  = multi_scalar_mul   src/curve.rs:141-158
//!     toy            the circuit of src/dvsnark_test.rs:34-128 proven as in :131-180:
//!                    { trapdoor[3], public[2], private[5], g_m, g_q, g_k_0, g_k_1, g_k_2 (file payloads after the
//!                      u64 count, hex), z_vals2inv, bar_wts (hex payloads), proof_bits (944 chars '0'/'1' =
//!                      Proof::to_bits), verify: true }
//!   reference_tree2n.bin   the `tree2n` file the toy setup wrote (src/tree_io.rs:144-214), byte for byte
//!
//! `multi_scalar_mul`, `point_scalar_mul(_gen)` and `CurvePoint::generator` are pub(crate) upstream, so the three
//! curve sections call the very xs233-sys functions those wrappers call, with the scalar bytes produced the way
//! `fr_to_le_bytes` (src/curve.rs:162-182) produces them.  Everything in `toy` goes through the public API only.
use ark_ff::{BigInteger, PrimeField, UniformRand};
use ark_std::rand::SeedableRng;
use dv_pari::curve::{CurvePoint, Fr};
use dv_pari::proving::{prover_prepares_precomputes, Proof};
use dv_pari::srs::{Trapdoor, SRS};
use rand_chacha::ChaCha20Rng;
use std::ffi::c_void;
use std::fmt::Write as _;
use std::io::Write as _;
use std::path::{Path, PathBuf};
use std::str::FromStr;
use xs233_sys::{xsk233_add, xsk233_encode, xsk233_generator, xsk233_mul_frob, xsk233_mulgen_frob, xsk233_neutral, xsk233_point};

fn hex(b: &[u8]) -> String {
    let mut s = String::with_capacity(2 * b.len());
    for x in b {
        write!(s, "{:02x}", x).unwrap();
    }
    s
}

/// canonical integer of an Fr as 0x-prefixed hex (what the Python side parses with int(.., 16))
fn fr_hex(x: &Fr) -> String {
    let h = hex(&x.into_bigint().to_bytes_be());
    let t = h.trim_start_matches('0');
    format!("0x{}", if t.is_empty() { "0" } else { t })
}

/// src/curve.rs:162-182, restated because it is private upstream
fn fr_to_le_bytes(fr: &Fr) -> Vec<u8> {
    let limbs = fr.into_bigint().0;
    let mut bytes = Vec::with_capacity(32);
    for limb in limbs.iter() {
        bytes.extend_from_slice(&limb.to_le_bytes());
    }
    bytes.truncate(30);
    while let Some(&last) = bytes.last() {
        if last == 0 {
            bytes.pop();
        } else {
            break;
        }
    }
    bytes
}

fn encode(p: &xsk233_point) -> [u8; 30] {
    let mut dst = [0u8; 30];
    unsafe { xsk233_encode(dst.as_mut_ptr() as *mut c_void, p) };
    dst
}
/// point_scalar_mul_gen, src/curve.rs:129-137
fn mulgen(k: &Fr) -> xsk233_point {
    let s = fr_to_le_bytes(k);
    unsafe {
        let mut r = xsk233_neutral;
        xsk233_mulgen_frob(&mut r, s.as_ptr() as *const _, s.len());
        r
    }
}
/// point_scalar_mul, src/curve.rs:113-126
fn mul(k: &Fr, p: &xsk233_point) -> xsk233_point {
    let s = fr_to_le_bytes(k);
    unsafe {
        let mut r = xsk233_neutral;
        xsk233_mul_frob(&mut r, p, s.as_ptr() as *const _, s.len());
        r
    }
}
fn add(a: &xsk233_point, b: &xsk233_point) -> xsk233_point {
    unsafe {
        let mut r = xsk233_neutral;
        xsk233_add(&mut r, a, b);
        r
    }
}

/// payload of an io_utils vector file: everything after the u64 little-endian count (src/io_utils.rs:1-7)
fn vec_file_payload(path: &Path) -> String {
    let b = std::fs::read(path).unwrap_or_else(|e| panic!("reading {}: {e}", path.display()));
    hex(&b[8..])
}

/// the dump of create_five_constraint_dump (src/dvsnark_test.rs:34-128) in the format of src/gnark_r1cs.rs:1-20
fn write_toy_r1cs(path: &Path) {
    let mut f = std::fs::File::create(path).unwrap();
    let be32 = |v: u64| {
        let mut o = [0u8; 32];
        o[24..].copy_from_slice(&v.to_be_bytes());
        o
    };
    let coeffs = [be32(1), be32(2)];
    f.write_all(&(coeffs.len() as u32).to_le_bytes()).unwrap();
    for c in coeffs.iter() {
        f.write_all(c).unwrap();
    }
    // wires: 0 one, 1 o, 2 w, 3 y, 4 z, 5 x, 6 t, 7 s;  (wire, coeff id)
    type T = (u32, u32);
    let rows: Vec<(Vec<T>, Vec<T>, Vec<T>)> = vec![
        (vec![(5, 0)], vec![(5, 0)], vec![(3, 0)]),
        (vec![(3, 0), (4, 0)], vec![(0, 0)], vec![(2, 0)]),
        (vec![(4, 1)], vec![(0, 0)], vec![(6, 0)]),
        (vec![(5, 0), (6, 0)], vec![(0, 0)], vec![(7, 0)]),
        (vec![(2, 0), (7, 0)], vec![(0, 0)], vec![(1, 0)]),
    ];
    f.write_all(&(rows.len() as u32).to_le_bytes()).unwrap();
    for (l, r, o) in rows.iter() {
        for n in [l.len(), r.len(), o.len()] {
            f.write_all(&(n as u32).to_le_bytes()).unwrap();
        }
        for t in l.iter().chain(r.iter()).chain(o.iter()) {
            f.write_all(&t.0.to_le_bytes()).unwrap();
            f.write_all(&t.1.to_le_bytes()).unwrap();
        }
    }
}

fn main() {
    let out_dir = PathBuf::from(std::env::args().nth(1).unwrap_or_else(|| "tests/golden".to_string()));
    std::fs::create_dir_all(&out_dir).unwrap();
    let mut j = String::from("{\n");
    writeln!(j, "  \"note\": \"captured from alpenlabs/dv-pari + xs233-sys =0.2.0 by rust-shim/examples/capture_golden.rs\",").unwrap();

    // ---- generator, neutral, k*G
    let g = unsafe { xsk233_generator };
    writeln!(j, "  \"generator\": \"{}\",", hex(&encode(&g))).unwrap();
    writeln!(j, "  \"generator_via_api\": \"{}\",", hex(&CurvePoint(g).to_bytes())).unwrap();
    writeln!(j, "  \"neutral\": \"{}\",", hex(&encode(&unsafe { xsk233_neutral }))).unwrap();
    let p_minus_1 = -Fr::from(1u64);
    let ks: Vec<Fr> = vec![
        Fr::from(0u64),
        Fr::from(1u64),
        Fr::from(2u64),
        Fr::from(3u64),
        Fr::from(255u64),
        Fr::from(256u64),
        Fr::from(0xdeadbeefu64),
        Fr::from(0xdeadbeefcafeu64),
        p_minus_1,
        Fr::from_str("2046321539021430469222588320254354836073765442868932981289742168423257").unwrap(), // 0x4be6fc..fb59
        Fr::from_str("1122341903156232028213042582636746400221451984116080623239229473258782").unwrap(), // 0x29a146..7d1e
    ];
    j.push_str("  \"mulgen\": [\n");
    for (i, k) in ks.iter().enumerate() {
        let e = encode(&mulgen(k));
        writeln!(j, "    {{\"k\": \"{}\", \"enc\": \"{}\"}}{}", fr_hex(k), hex(&e), if i + 1 < ks.len() { "," } else { "" }).unwrap();
    }
    j.push_str("  ],\n");

    // ---- k * P for non-generator bases, and a 16-term multi_scalar_mul (per-point products, then the sum)
    let mut rng = ChaCha20Rng::seed_from_u64(0xD5A1_0016);
    let bases: Vec<xsk233_point> = (0..16).map(|_| mulgen(&Fr::rand(&mut rng))).collect();
    let scalars: Vec<Fr> = (0..16).map(|_| Fr::rand(&mut rng)).collect();
    j.push_str("  \"mul\": [\n");
    for i in 0..4 {
        let e = encode(&mul(&scalars[i], &bases[i]));
        writeln!(j, "    {{\"k\": \"{}\", \"base\": \"{}\", \"enc\": \"{}\"}}{}", fr_hex(&scalars[i]), hex(&encode(&bases[i])), hex(&e),
                 if i < 3 { "," } else { "" }).unwrap();
    }
    j.push_str("  ],\n");
    let mut acc = unsafe { xsk233_neutral };
    for i in 0..16 {
        acc = add(&acc, &mul(&scalars[i], &bases[i]));
    }
    j.push_str("  \"msm16\": {\n    \"scalars\": [");
    j.push_str(&scalars.iter().map(|s| format!("\"{}\"", fr_hex(s))).collect::<Vec<_>>().join(", "));
    j.push_str("],\n    \"points\": [");
    j.push_str(&bases.iter().map(|p| format!("\"{}\"", hex(&encode(p)))).collect::<Vec<_>>().join(", "));
    writeln!(j, "],\n    \"result\": \"{}\"\n  }},", hex(&encode(&acc))).unwrap();

    // ---- the toy circuit end to end through the public API (src/dvsnark_test.rs:131-180)
    let cache = std::env::temp_dir().join("dvpari_capture_golden_cache");
    let _ = std::fs::remove_dir_all(&cache);
    std::fs::create_dir_all(&cache).unwrap();
    write_toy_r1cs(&cache.join(dv_pari::artifacts::R1CS_CONSTRAINTS_FILE));
    let x = Fr::from(3u64);
    let y = x * x;
    let z = Fr::from(4u64);
    let w = y + z;
    let t = z + z;
    let s = x + t;
    let o = w + s;
    let public_inputs: Vec<Fr> = vec![o, w];
    let witness = vec![y, z, x, t, s];
    let mut rng = ChaCha20Rng::seed_from_u64(43);
    let trapdoor = Trapdoor { tau: Fr::rand(&mut rng), delta: Fr::rand(&mut rng), epsilon: Fr::rand(&mut rng) };
    let _ = SRS::verifier_runs_setup(trapdoor, &cache, public_inputs.len(), true, true).unwrap();
    prover_prepares_precomputes(&cache, true).unwrap();
    let proof = Proof::prove(cache.to_str().unwrap(), public_inputs.clone(), &witness);
    let ok = SRS::verify(trapdoor, &public_inputs, &proof);
    assert!(ok, "the reference rejected its own proof");
    let bits: String = proof.to_bits().iter().map(|b| if *b { '1' } else { '0' }).collect();
    j.push_str("  \"toy\": {\n");
    writeln!(j, "    \"trapdoor\": [\"{}\", \"{}\", \"{}\"],", fr_hex(&trapdoor.tau), fr_hex(&trapdoor.delta), fr_hex(&trapdoor.epsilon)).unwrap();
    writeln!(j, "    \"public\": [{}],", public_inputs.iter().map(|v| format!("\"{}\"", fr_hex(v))).collect::<Vec<_>>().join(", ")).unwrap();
    writeln!(j, "    \"private\": [{}],", witness.iter().map(|v| format!("\"{}\"", fr_hex(v))).collect::<Vec<_>>().join(", ")).unwrap();
    for name in ["g_m", "g_q", "g_k_0", "g_k_1", "g_k_2", "z_vals2inv", "bar_wts"] {
        writeln!(j, "    \"{}\": \"{}\",", name, vec_file_payload(&cache.join(name))).unwrap();
    }
    writeln!(j, "    \"commit_p\": \"{}\",", hex(&proof.commit_p)).unwrap();
    writeln!(j, "    \"kzg_k\": \"{}\",", hex(&proof.kzg_k)).unwrap();
    writeln!(j, "    \"proof_bits\": \"{}\",", bits).unwrap();
    writeln!(j, "    \"verify\": {}", ok).unwrap();
    j.push_str("  }\n}\n");
    std::fs::write(out_dir.join("reference_v1.json"), j).unwrap();
    // the reference's own tree file (16 leaves: D u D' of the 8-row padded toy circuit)
    std::fs::copy(cache.join("tree2n"), out_dir.join("reference_tree2n.bin")).unwrap();
    println!("wrote {}/reference_v1.json and reference_tree2n.bin", out_dir.display());
}
