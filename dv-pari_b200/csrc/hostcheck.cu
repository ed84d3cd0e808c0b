// dvp_hostcheck_op: runs the SAME __host__ __device__ source that the kernels execute, on the CPU.
// It exists only so the `-m "not gpu"` tests can pin the device arithmetic (GF(2^233), Fr, point
// addition, the 30-byte codec, the PCLMUL host tail) against the oracle without a GPU.  No product
// path calls it; every dvp_* entry point that computes needs a CUDA device.
#include <cstring>
#include "../../include/dvpari.h"
#include "fr.cuh"
#include "fr29.cuh"
#include "host_gf.hpp"
#include "k233_codec.cuh"
#include "k233_ld.cuh"

using namespace dvp;

extern "C" int dvp_hostcheck_op(int op, const void *a_, const void *b_, void *out_, size_t n) {
    const uint8_t *a = (const uint8_t *)a_, *b = (const uint8_t *)b_;
    uint8_t *out = (uint8_t *)out_;
    for (size_t i = 0; i < n; i++) {
        switch (op) {
        case 0: case 1: case 2: case 6: case 7: {
            gf x, y = gf_zero(), r;
            memcpy(x.v, a + i * 32, 32);
            if (b) memcpy(y.v, b + i * 32, 32);
            r = op == 0 ? gf_mul(x, y) : op == 1 ? gf_sqr(x) : op == 2 ? gf_inv(x) : op == 6 ? host::hmul(x, y) : host::hinv(x);
            memcpy(out + i * 32, r.v, 32);
            break;
        }
        case 3: {
            fr x, y;
            memcpy(x.v, a + i * 32, 32);
            memcpy(y.v, b + i * 32, 32);
            fr r = fr_mul(x, y);
            memcpy(out + i * 32, r.v, 32);
            break;
        }
        case 4: {
            fr x;
            memcpy(x.v, a + i * 32, 32);
            uint32_t c[8];
            fr_to_canonical(c, x);
            memcpy(out + i * 32, c, 32);
            break;
        }
        case 5: { // affine add, 64-byte points
            AffPt p, q;
            memcpy(&p, a + i * 64, 64);
            memcpy(&q, b + i * 64, 64);
            AffPt r = pt_add_slow(p, q);
            memcpy(out + i * 64, &r, 64);
            break;
        }
        case 8: { // decode: 30 bytes -> 64-byte point + validity byte (out stride 65)
            AffPt p;
            bool ok = xsk233_decode_pt(a + i * 30, p);
            memcpy(out + i * 65, &p, 64);
            out[i * 65 + 64] = ok ? 1 : 0;
            break;
        }
        case 9: { // encode: 64-byte point -> 30 bytes (device formula) ; 10: host tail formula
            AffPt p;
            memcpy(&p, a + i * 64, 64);
            xsk233_encode_pt(out + i * 30, p);
            break;
        }
        case 10: {
            AffPt p;
            memcpy(&p, a + i * 64, 64);
            host::encode30(out + i * 30, p);
            break;
        }
        case 11: { // host LD accumulator: 2*a + b  (a, b affine 64-byte) -> affine
            AffPt p, q;
            memcpy(&p, a + i * 64, 64);
            memcpy(&q, b + i * 64, 64);
            host::LdPt acc = host::ld_add_affine(host::ld_inf(), p);
            acc = host::ld_dbl(acc);
            acc = host::ld_add_affine(acc, q);
            AffPt r = host::ld_to_affine(acc);
            memcpy(out + i * 64, &r, 64);
            break;
        }
        case 12: { // fr_add / 13 fr_sub / 14 fr_inv
            fr x, y;
            memcpy(x.v, a + i * 32, 32);
            memcpy(y.v, b + i * 32, 32);
            fr r = fr_add(x, y);
            memcpy(out + i * 32, r.v, 32);
            break;
        }
        case 13: {
            fr x, y;
            memcpy(x.v, a + i * 32, 32);
            memcpy(y.v, b + i * 32, 32);
            fr r = fr_sub(x, y);
            memcpy(out + i * 32, r.v, 32);
            break;
        }
        case 14: {
            fr x;
            memcpy(x.v, a + i * 32, 32);
            fr r = fr_inv(x);
            memcpy(out + i * 32, r.v, 32);
            break;
        }
        case 15: { // ECFFT butterfly row on 29-bit limbs: a = (m0, x0), b = (m1, x1), 64 bytes each -> m0 x0 + m1 x1
            fr m0, x0, m1, x1;
            memcpy(m0.v, a + i * 64, 32);
            memcpy(x0.v, a + i * 64 + 32, 32);
            memcpy(m1.v, b + i * 64, 32);
            memcpy(x1.v, b + i * 64 + 32, 32);
            const fr29 r = fr29_dot2(fr29_prescale(m0), fr29_from_fr(x0), fr29_prescale(m1), fr29_from_fr(x1));
            const fr o = fr_from_fr29(r);
            memcpy(out + i * 32, o.v, 32);
            break;
        }
        case 19: { // normalised butterfly row on 29-bit limbs: a = (m, x0), b = (unused, x1) -> x0 + m x1
            fr m, x0, x1;
            memcpy(m.v, a + i * 64, 32);
            memcpy(x0.v, a + i * 64 + 32, 32);
            memcpy(x1.v, b + i * 64 + 32, 32);
            const fr o = fr_from_fr29(fr29_muladd(fr29_prescale(m), fr29_from_fr(x1), fr29_from_fr(x0)));
            memcpy(out + i * 32, o.v, 32);
            break;
        }
        case 20: case 21: { // semi-reduced rows on raw 232-bit operands: a = (m, x0), b = (m1, x1), m and m1 Montgomery < p;
            // 20: x0 + m x1, 21: m x0 + m1 x1 -> 8 x 32-bit words of the semi-reduced result
            fr m, x0, m1, x1;
            memcpy(m.v, a + i * 64, 32);
            memcpy(x0.v, a + i * 64 + 32, 32);
            memcpy(m1.v, b + i * 64, 32);
            memcpy(x1.v, b + i * 64 + 32, 32);
            const fr29 r = op == 20 ? fr29_muladd_semi(fr29_prescale(m), fr29_from_fr(x1), fr29_from_fr(x0))
                                    : fr29_dot2_semi(fr29_prescale(m), fr29_from_fr(x0), fr29_prescale(m1), fr29_from_fr(x1));
            uint32_t ok = 1;
            for (int q = 0; q < 8; q++) ok &= r.l[q] <= DVP_M29;
            fr o = fr_from_fr29(r);
            if (!ok) memset(o.v, 0xff, 32); // limbs not normalised
            memcpy(out + i * 32, o.v, 32);
            break;
        }
        case 18: { // three-term dot product on 29-bit limbs + fr29_add: a = (m0, x0), b = (m1, x1) -> m0 x0 + m1 x1 + m0 x1, then + m1
            fr m0, x0, m1, x1;
            memcpy(m0.v, a + i * 64, 32);
            memcpy(x0.v, a + i * 64 + 32, 32);
            memcpy(m1.v, b + i * 64, 32);
            memcpy(x1.v, b + i * 64 + 32, 32);
            const fr29 mm[3] = {fr29_prescale(m0), fr29_prescale(m1), fr29_prescale(m0)};
            const fr29 xx[3] = {fr29_from_fr(x0), fr29_from_fr(x1), fr29_from_fr(x1)};
            const fr29 r = fr29_add(fr29_dotn<3>(mm, xx), fr29_from_fr(m1));
            const fr o = fr_from_fr29(r);
            memcpy(out + i * 32, o.v, 32);
            break;
        }
        case 16: { // complete LD addition on projective operands: (2a) + (b + a) = 3a + b
            AffPt p, q;
            memcpy(&p, a + i * 64, 64);
            memcpy(&q, b + i * 64, 64);
            const LdPt P1 = ld_double(ld_from_affine(p));
            const LdPt P2 = ld_add(ld_from_affine(q), ld_from_affine(p));
            const LdPt S = ld_add(P1, P2);
            const AffPt r = ld_to_affine_with(S, gf_inv(S.Z));
            memcpy(out + i * 64, &r, 64);
            break;
        }
        case 17: { // LD addition of two projective copies of the same operands: (a + b) + (b + a) and (a + b) - (b + a)
            AffPt p, q;
            memcpy(&p, a + i * 64, 64);
            memcpy(&q, b + i * 64, 64);
            const LdPt P1 = ld_add(ld_from_affine(p), ld_from_affine(q));
            LdPt P2 = ld_add(ld_double(ld_from_affine(q)), ld_add(ld_from_affine(p), ld_from_affine(pt_neg(q)))); // = a + b, other Z
            const LdPt S = ld_add(P1, P2); // doubling branch
            const AffPt r = ld_to_affine_with(S, gf_inv(S.Z));
            memcpy(out + i * 128, &r, 64);
            // negate P2: y -> y + x, in LD: Y -> Y + X Z
            P2.Y = gf_add(P2.Y, gf_mul(P2.X, P2.Z));
            const LdPt D = ld_add(P1, P2); // opposite branch -> infinity
            const AffPt r2 = ld_to_affine_with(D, gf_inv(D.Z));
            memcpy(out + i * 128 + 64, &r2, 64);
            break;
        }
        default:
            return DVP_ERR_BAD_ARG;
        }
    }
    return DVP_OK;
}
