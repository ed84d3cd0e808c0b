/*
 * ORACLE CROSS-CHECK (test infrastructure only): thin wrappers over OpenSSL libcrypto
 * (BN_GF2m_* with {233,74,0} and EC_POINT_* on NID_sect233k1) used by tests/ to pin
 * gf233.c and k233.c against an independent implementation.  Big-endian 30-byte values.
 */
#include <openssl/bn.h>
#include <openssl/ec.h>
#include <openssl/obj_mac.h>
#include <string.h>

static const int POLY[] = {233, 74, 0, -1};

static void put30(unsigned char out[30], const BIGNUM *a) {
    memset(out, 0, 30);
    BN_bn2binpad(a, out, 30);
}
/* op: 0 mul, 1 sqr(a), 2 inv(a), 3 sqrt(a); returns 1 on success */
int ossl_gf_op(int op, const unsigned char a[30], const unsigned char b[30], unsigned char out[30]) {
    BN_CTX *ctx = BN_CTX_new();
    BIGNUM *A = BN_bin2bn(a, 30, NULL), *B = BN_bin2bn(b, 30, NULL), *R = BN_new();
    int ok = 0;
    if (op == 0) ok = BN_GF2m_mod_mul_arr(R, A, B, POLY, ctx);
    else if (op == 1) ok = BN_GF2m_mod_sqr_arr(R, A, POLY, ctx);
    else if (op == 2) ok = BN_GF2m_mod_inv_arr(R, A, POLY, ctx);
    else if (op == 3) ok = BN_GF2m_mod_sqrt_arr(R, A, POLY, ctx);
    if (ok) put30(out, R);
    BN_free(A); BN_free(B); BN_free(R); BN_CTX_free(ctx);
    return ok;
}
/* out = k*P (+ Q if q_x != NULL) on sect233k1; k big-endian klen bytes; returns 1 ok, 2 infinity, 0 error */
int ossl_ec_mul_add(const unsigned char *k, int klen, const unsigned char px[30], const unsigned char py[30],
                    const unsigned char *qx, const unsigned char *qy, unsigned char ox[30], unsigned char oy[30]) {
    EC_GROUP *g = EC_GROUP_new_by_curve_name(NID_sect233k1);
    BN_CTX *ctx = BN_CTX_new();
    EC_POINT *P = EC_POINT_new(g), *R = EC_POINT_new(g);
    BIGNUM *K = BN_bin2bn(k, klen, NULL), *X = BN_bin2bn(px, 30, NULL), *Y = BN_bin2bn(py, 30, NULL);
    int ret = 0;
    if (!EC_POINT_set_affine_coordinates(g, P, X, Y, ctx)) goto done;
    if (!EC_POINT_mul(g, R, NULL, P, K, ctx)) goto done;
    if (qx) {
        EC_POINT *Q = EC_POINT_new(g);
        BIGNUM *QX = BN_bin2bn(qx, 30, NULL), *QY = BN_bin2bn(qy, 30, NULL);
        int ok = EC_POINT_set_affine_coordinates(g, Q, QX, QY, ctx) && EC_POINT_add(g, R, R, Q, ctx);
        BN_free(QX); BN_free(QY); EC_POINT_free(Q);
        if (!ok) goto done;
    }
    if (EC_POINT_is_at_infinity(g, R)) { ret = 2; goto done; }
    if (!EC_POINT_get_affine_coordinates(g, R, X, Y, ctx)) goto done;
    put30(ox, X); put30(oy, Y);
    ret = 1;
done:
    BN_free(K); BN_free(X); BN_free(Y); EC_POINT_free(P); EC_POINT_free(R); BN_CTX_free(ctx); EC_GROUP_free(g);
    return ret;
}
