/*
 * spinoza_oracle.c -- CPU restatement of QuState/spinoza's gate-application hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (spinoza_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED against the reference's own golden vectors (tests/test_oracle_golden.py
 * replays every known-answer test of spinoza/src/gates.rs:1531-2236, core.rs:271-339,
 * measurement.rs:100-246, circuit.rs:622-1250).  The real reference (Rust, nightly) cannot be
 * compiled in this image (no cargo/rustc), so cpu_baseline.kind is "port".
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/spinoza/src/).  Arithmetic mirrors the reference operation by operation:
 * plain `*`/`+`/`-` where the reference uses them, fma() where the reference calls
 * `mul_add`, so this file must be compiled with -ffp-contract=off (see oracle/Makefile).
 * The four reference defects listed in SURVEY.md 2.3 are NOT reproduced:
 *   B1  gates.rs:593   parallel H writes b1 + c1      -> serial formula gates.rs:551-560 used
 *   B2  gates.rs:401-446,807-827,912-928,1209-1229 scan-and-skip without target-bit check
 *                                                     -> intended pair set (control bits set, target bit 0)
 *       orc_mc_scan_literal() restates the literal loop (bounds-checked) so tests can show the two
 *       agree on the reference's safe domain.
 *   B3  gates.rs:1209-1229 ry_mc_apply applies RX     -> true RY (gates.rs:1116-1119)
 *   B4  gates.rs:298-311  zeros are dropped from the mask -> MIRRORED (it is API behaviour, not UB)
 *
 * Threading mirrors the reference's rayon decomposition (for the CPU baseline timing):
 *   - uncontrolled gates: parallel over the 2^(n-1-t) chunks, serial inner loop (gates.rs:361-372)
 *   - RZ: parallel over chunks of 2^t amplitudes (gates.rs:949-963)
 *   - single-controlled gates: parallel per pair index (gates.rs:392-397)
 *   - u_c_apply, cc_apply, mc_apply, swap_apply: serial (gates.rs:1356, 405, 429, 1379)
 *   - serial whenever threads < 2 or n < LOW_QUBIT_THRESHOLD = 15 (gates.rs:14,351)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* Gate kinds: same numbering as include/spinoza_b200.h (spz_gate_kind). gates.rs:44-74 */
enum {
    ORC_H = 0, ORC_M = 1, ORC_X = 2, ORC_Y = 3, ORC_Z = 4, ORC_P = 5, ORC_RX = 6, ORC_RY = 7,
    ORC_RZ = 8, ORC_SWAP = 9, ORC_U = 10, ORC_UNITARY = 11, ORC_BITFLIP = 12
};

enum { ORC_OK = 0, ORC_ERR_INVALID = 1, ORC_ERR_UNSUPPORTED = 2 };

#define LOW_QUBIT_THRESHOLD 15 /* gates.rs:14 */
static const double SQRT_ONE_HALF = 0.70710678118654752440; /* math.rs:5 FRAC_1_SQRT_2 */

static int g_threads = 1;

ORC_API void orc_set_threads(int t) {
    g_threads = t < 1 ? 1 : t;
#ifdef _OPENMP
    omp_set_num_threads(g_threads);
#endif
}
ORC_API int orc_get_threads(void) { return g_threads; }
ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

/* gates.rs:351 -- `Config::global().threads < 2 || state.n < LOW_QUBIT_THRESHOLD` */
static inline int serial_path(int n) { return g_threads < 2 || n < LOW_QUBIT_THRESHOLD; }

typedef int64_t idx_t;

/* ---- State::new  core.rs:32-42 ------------------------------------------------------------ */
ORC_API void orc_state_init(double *re, double *im, int n) {
    idx_t len = (idx_t)1 << n;
    memset(re, 0, (size_t)len * sizeof(double));
    memset(im, 0, (size_t)len * sizeof(double));
    re[0] = 1.0;
}

/* ---- per-pair updates -------------------------------------------------------------------- */

/* gates.rs:329-334 x_apply_target */
static inline void x_pair(double *re, double *im, idx_t s0, idx_t s1) {
    double t = re[s0]; re[s0] = re[s1]; re[s1] = t;
    t = im[s0]; im[s0] = im[s1]; im[s1] = t;
}
/* gates.rs:455-464: swap s0<->s1, swap re<->im at both, negate im(s0), re(s1)  => s0=(d,-c) s1=(-b,a) */
static inline void y_pair(double *re, double *im, idx_t s0, idx_t s1) {
    double a = re[s0], b = im[s0], c = re[s1], d = im[s1];
    re[s0] = d;  im[s0] = -c;
    re[s1] = -b; im[s1] = a;
}
/* gates.rs:543-561 h_apply_strat2 (serial formula; B1 not reproduced) */
static inline void h_pair(double *re, double *im, idx_t s0, idx_t s1) {
    double a = re[s0], b = im[s0], c = re[s1], d = im[s1];
    double a1 = SQRT_ONE_HALF * a, b1 = SQRT_ONE_HALF * b;
    double c1 = SQRT_ONE_HALF * c, d1 = SQRT_ONE_HALF * d;
    re[s0] = a1 + c1; im[s0] = b1 + d1;
    re[s1] = a1 - c1; im[s1] = b1 - d1;
}
/* gates.rs:703-726 rx_apply_target */
static inline void rx_pair(double *re, double *im, idx_t s0, idx_t s1, double cs, double neg_sin) {
    double a = re[s0], b = im[s0], c = re[s1], d = im[s1];
    re[s0] = a * cs - d * neg_sin;
    im[s0] = b * cs + c * neg_sin;
    re[s1] = b * -neg_sin + c * cs;
    im[s1] = d * cs + a * neg_sin;
}
/* gates.rs:1104-1120 ry_apply_strategy2 */
static inline void ry_pair(double *re, double *im, idx_t s0, idx_t s1, double sn, double cs) {
    double a = re[s0], b = im[s0], c = re[s1], d = im[s1];
    re[s0] = a * cs - c * sn;
    im[s0] = b * cs - d * sn;
    re[s1] = a * sn + c * cs;
    im[s1] = b * sn + d * cs;
}
/* gates.rs:834-839 p_proc_chunk: z *= cos + i sin via mul_add */
static inline void p_one(double *re, double *im, idx_t s1, double cs, double sn) {
    double zr = re[s1], zi = im[s1];
    re[s1] = fma(zr, cs, -zi * sn);
    im[s1] = fma(zi, cs, zr * sn);
}
/* gates.rs:940-946 rz element update: (c,d) *= m */
static inline void rz_one(double *re, double *im, idx_t i, double mre, double mim) {
    double c = re[i], d = im[i];
    re[i] = fma(c, mre, -d * mim);
    im[i] = fma(c, mim, d * mre);
}
/* gates.rs:1231-1267 u_apply_target.  g = {a, k, l, q, r, s, t}; g[0].im assumed 0 (gates.rs:1247) */
static inline void u_pair(double *re, double *im, idx_t s0, idx_t s1, const double *g) {
    double c = re[s0], d = im[s0], m = re[s1], n = im[s1];
    double a = g[0], k = g[1], l = g[2], q = g[3], r = g[4], s = g[5], t = g[6];
    double t0 = fma(a, c, fma(k, m, -l * n));
    double t1 = fma(a, d, fma(k, n, l * m));
    double t2 = fma(q, c, fma(-r, d, fma(s, m, -t * n)));
    double t3 = fma(q, d, fma(r, c, fma(s, n, t * m)));
    re[s0] = t0; im[s0] = t1; re[s1] = t2; im[s1] = t3;
}
/* gates.rs:1286-1304 (u_apply) == gates.rs:1330-1348 (u_c_apply): host scalars of U(theta,phi,lambda) */
ORC_API void orc_u_scalars(double theta, double phi, double lambda, double *g) {
    double st = sin(theta * 0.5), ct = cos(theta * 0.5);
    double sl = sin(lambda), cl = cos(lambda);
    double spl = sin(phi + lambda), cpl = cos(phi + lambda);
    double sp = sin(phi), cp = cos(phi);
    g[0] = ct;            /* c.re */
    g[1] = -cl * st;      /* ncs.re */
    g[2] = -sl * st;      /* ncs.im */
    g[3] = cp * st;       /* es.re */
    g[4] = sp * st;       /* es.im */
    g[5] = cpl * ct;      /* ec.re */
    g[6] = spl * ct;      /* ec.im */
}

/* Resolved scalars of one gate, computed once on the host exactly as each *_apply does. */
typedef struct { int kind; double s[7]; } gate_scalars;

static int resolve(int kind, const double *p, gate_scalars *gs) {
    gs->kind = kind;
    memset(gs->s, 0, sizeof gs->s);
    switch (kind) {
    case ORC_H: case ORC_X: case ORC_Y: case ORC_Z: return ORC_OK;
    case ORC_P: /* gates.rs:865 sin_cos(angle) */
        gs->s[0] = cos(p[0]); gs->s[1] = sin(p[0]); return ORC_OK;
    case ORC_RX: { /* gates.rs:749-751 */
        double th = p[0] * 0.5; gs->s[0] = cos(th); gs->s[1] = -sin(th); return ORC_OK; }
    case ORC_RY: { /* gates.rs:1125-1126: (sin, cos) */
        double th = p[0] * 0.5; gs->s[0] = sin(th); gs->s[1] = cos(th); return ORC_OK; }
    case ORC_RZ: { /* gates.rs:970-973: d0=(c,-s) d1=(c,s) */
        double th = p[0] * 0.5; gs->s[0] = cos(th); gs->s[1] = sin(th); return ORC_OK; }
    case ORC_U: orc_u_scalars(p[0], p[1], p[2], gs->s); return ORC_OK;
    default: return ORC_ERR_UNSUPPORTED;
    }
}

/* Apply the resolved gate to the pair (s0,s1).  Z/P only touch s1; RZ touches both. */
static inline void pair_update(const gate_scalars *g, double *re, double *im, idx_t s0, idx_t s1) {
    switch (g->kind) {
    case ORC_H: h_pair(re, im, s0, s1); break;
    case ORC_X: x_pair(re, im, s0, s1); break;
    case ORC_Y: y_pair(re, im, s0, s1); break;
    case ORC_Z: re[s1] = -re[s1]; im[s1] = -im[s1]; break;               /* gates.rs:1051-1055 */
    case ORC_P: p_one(re, im, s1, g->s[0], g->s[1]); break;
    case ORC_RX: rx_pair(re, im, s0, s1, g->s[0], g->s[1]); break;
    case ORC_RY: ry_pair(re, im, s0, s1, g->s[0], g->s[1]); break;
    case ORC_RZ: rz_one(re, im, s0, g->s[0], -g->s[1]);                   /* d0, gates.rs:998-1007 */
                 rz_one(re, im, s1, g->s[0], g->s[1]); break;             /* d1, gates.rs:1008-1017 */
    case ORC_U: u_pair(re, im, s0, s1, g->s); break;
    default: break;
    }
}

/* ---- swap_apply  gates.rs:1376-1386 (serial scan) ---------------------------------------- */
ORC_API int orc_swap(double *re, double *im, int n, int t0, int t1) {
    if (t0 < 0 || t1 < 0 || t0 >= n || t1 >= n) return ORC_ERR_INVALID; /* assert gates.rs:1377 */
    idx_t len = (idx_t)1 << n;
    for (idx_t i = 0; i < len; ++i) {
        if (((i >> t0) & 1) == 0 && ((i >> t1) & 1) == 1) {
            idx_t j = i + ((idx_t)1 << t0) - ((idx_t)1 << t1);
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    return ORC_OK;
}

/* ---- apply  gates.rs:215-232 -------------------------------------------------------------- */
ORC_API int orc_apply(int kind, const double *params, int t0, int t1, double *re, double *im, int n,
                      int target) {
    if (kind == ORC_SWAP) return orc_swap(re, im, n, t0, t1);               /* gates.rs:225 */
    if (target < 0 || target >= n) return ORC_ERR_INVALID;
    gate_scalars g;
    int rc = resolve(kind, params, &g);
    if (rc) return rc;                                                       /* gates.rs:230 unimplemented!() */
    const idx_t len = (idx_t)1 << n;
    const int ser = serial_path(n);

    if (kind == ORC_RZ) {
        /* rz_apply_strategy1 gates.rs:930-965: chunks of 2^t, factor diag[i & 1] */
        const idx_t chunk = (idx_t)1 << target, nchunks = len >> target;
        const double c = g.s[0], s = g.s[1];
        #pragma omp parallel for schedule(static) if (!ser)
        for (idx_t ci = 0; ci < nchunks; ++ci) {
            const double mim = (ci & 1) ? s : -s;
            double *r = re + ci * chunk, *m = im + ci * chunk;
            for (idx_t j = 0; j < chunk; ++j) rz_one(r, m, j, c, mim);
        }
        return ORC_OK;
    }
    /* every other gate: chunks = 2^(n-1-t), base = (2*chunk) << t, serial inner loop over dist
       (x: gates.rs:336-374, y: 448-484, h: 536-616, rx: 728-776, p: 829-880, z: 1048-1088,
        ry: 1090-1142, u: 1269-1320).  The target==0 special cases (gates.rs:350-359, 753-762)
       visit the same pairs with the same arithmetic. */
    const idx_t dist = (idx_t)1 << target, chunks = (len >> 1) >> target;
    #pragma omp parallel for schedule(static) if (!ser)
    for (idx_t ch = 0; ch < chunks; ++ch) {
        const idx_t base = (2 * ch) << target;
        for (idx_t i = 0; i < dist; ++i) pair_update(&g, re, im, base + i, base + i + dist);
    }
    return ORC_OK;
}

/* ---- c_apply  gates.rs:257-269 ------------------------------------------------------------ */
ORC_API int orc_c_apply(int kind, const double *params, double *re, double *im, int n, int control,
                        int target) {
    if (control < 0 || target < 0 || control >= n || target >= n || control == target)
        return ORC_ERR_INVALID;
    /* supported set gates.rs:258-267: H X Y P RX RY RZ U ; everything else is todo!() */
    if (!(kind == ORC_H || kind == ORC_X || kind == ORC_Y || kind == ORC_P || kind == ORC_RX ||
          kind == ORC_RY || kind == ORC_RZ || kind == ORC_U))
        return ORC_ERR_UNSUPPORTED;
    gate_scalars g;
    int rc = resolve(kind, params, &g);
    if (rc) return rc;
    const idx_t len = (idx_t)1 << n, dist = (idx_t)1 << target;

    if (kind == ORC_U) {
        /* u_c_apply gates.rs:1353-1362: serial scan, i & mask == mask -> s1 = i */
        const idx_t mask = ((idx_t)1 << control) | ((idx_t)1 << target);
        for (idx_t i = 0; i < len; ++i)
            if ((i & mask) == mask) u_pair(re, im, i - dist, i, g.s);
        return ORC_OK;
    }
    /* gates.rs:380-398 (and twins): end = len>>2, marks=(min,max),
       x = i + (1<<(M-1)) + ((i>>(M-1))<<(M-1));  s1 = x + (1<<m) + ((x>>m)<<m);  s0 = s1 - dist */
    const idx_t end = len >> 2;
    const int m0 = control < target ? control : target, m1 = control < target ? target : control;
    const int ser = serial_path(n);
    #pragma omp parallel for schedule(static) if (!ser)
    for (idx_t i = 0; i < end; ++i) {
        idx_t x = i + ((idx_t)1 << (m1 - 1)) + ((i >> (m1 - 1)) << (m1 - 1));
        idx_t s1 = x + ((idx_t)1 << m0) + ((x >> m0) << m0);
        idx_t s0 = s1 - dist;
        pair_update(&g, re, im, s0, s1);
    }
    return ORC_OK;
}

/* ---- mc_apply mask  gates.rs:298-311 (B4 mirrored: zeros are dropped from the mask) -------- */
ORC_API uint64_t orc_mc_mask(const int *controls, int nc, const int *zeros, int nz) {
    uint64_t mask = 0;
    for (int i = 0; i < nc; ++i) {
        int skip = 0;
        for (int j = 0; j < nz; ++j) if (zeros[j] == controls[i]) skip = 1;
        if (!skip) mask |= (uint64_t)1 << controls[i];
    }
    return mask;
}

/* Intended semantics of x_cc_apply / x_mc_apply / p_mc_apply / rx_mc_apply / ry_mc_apply
   (gates.rs:401-446, 807-827, 912-928, 1209-1229): every pair (s0, s0+dist) with target bit of
   s0 clear and (s0 & mask) == mask.  B2 (missing target-bit check) and B3 (RY->RX) not reproduced. */
ORC_API int orc_mc_apply_mask(int kind, const double *params, double *re, double *im, int n,
                              uint64_t mask, int target) {
    if (target < 0 || target >= n) return ORC_ERR_INVALID;
    if ((mask >> target) & 1) return ORC_ERR_INVALID;
    if (n < 64 && (mask >> n)) return ORC_ERR_INVALID;
    /* supported set gates.rs:313-319: X P RX RY */
    if (!(kind == ORC_X || kind == ORC_P || kind == ORC_RX || kind == ORC_RY))
        return ORC_ERR_UNSUPPORTED;
    gate_scalars g;
    int rc = resolve(kind, params, &g);
    if (rc) return rc;
    const idx_t len = (idx_t)1 << n, dist = (idx_t)1 << target;
    for (idx_t i = 0; i < len; ++i)
        if (((uint64_t)i & mask) == mask && !(i & dist)) pair_update(&g, re, im, i, i + dist);
    return ORC_OK;
}

ORC_API int orc_mc_apply(int kind, const double *params, double *re, double *im, int n,
                         const int *controls, int nc, const int *zeros, int nz, int target) {
    if (!(n > nc)) return ORC_ERR_INVALID; /* debug_assert gates.rs:297 */
    for (int i = 0; i < nc; ++i)
        if (controls[i] < 0 || controls[i] >= n || controls[i] == target) return ORC_ERR_INVALID;
    return orc_mc_apply_mask(kind, params, re, im, n, orc_mc_mask(controls, nc, zeros, nz), target);
}

/* cc_apply gates.rs:272-277: X only */
ORC_API int orc_cc_apply(int kind, const double *params, double *re, double *im, int n, int c0, int c1,
                         int target) {
    if (kind != ORC_X) return ORC_ERR_UNSUPPORTED;
    if (c0 < 0 || c1 < 0 || c0 >= n || c1 >= n || c0 == target || c1 == target) return ORC_ERR_INVALID;
    return orc_mc_apply_mask(kind, params, re, im, n, ((uint64_t)1 << c0) | ((uint64_t)1 << c1), target);
}

/* LITERAL restatement of the reference's scan-and-skip loop (gates.rs:425-446 and twins), with the
   out-of-bounds access turned into an error instead of UB.  Returns ORC_ERR_INVALID when the literal
   loop would leave the state (i.e. outside B2's safe domain).  kind: X, P, RX, or RY (RY runs the RX
   update with (cos,-sin), exactly as gates.rs:1209-1229 does -- defect B3 -- so tests can show it). */
ORC_API int orc_mc_scan_literal(int kind, const double *params, double *re, double *im, int n,
                                uint64_t mask, int target) {
    const idx_t len = (idx_t)1 << n, dist = (idx_t)1 << target;
    double th = params ? params[0] : 0.0;
    double pc = cos(th), ps = sin(th);                 /* p_mc_apply gates.rs:913 */
    double ct = cos(th * 0.5), nst = -sin(th * 0.5);   /* rx_mc_apply gates.rs:811-813 */
    idx_t i = 0;
    while (i < len) {
        if (((uint64_t)i & mask) == mask) {
            idx_t s0 = i, s1 = i + dist;
            if (s1 >= len) return ORC_ERR_INVALID;
            if (kind == ORC_X) x_pair(re, im, s0, s1);
            else if (kind == ORC_P) p_one(re, im, s1, pc, ps);
            else if (kind == ORC_RX || kind == ORC_RY) rx_pair(re, im, s0, s1, ct, nst);
            else return ORC_ERR_UNSUPPORTED;
            i += dist;
        }
        i += 1;
    }
    return ORC_OK;
}

/* ---- iqft  core.rs:184-191 ----------------------------------------------------------------- */
static double pow2f(int e) { return ldexp(1.0, e); } /* math.rs:33-36 */

ORC_API int orc_iqft(double *re, double *im, int n, const int *targets, int m) {
    const double PI = 3.14159265358979323846;
    for (int j = m - 1; j >= 0; --j) {
        int rc = orc_apply(ORC_H, NULL, 0, 0, re, im, n, targets[j]);
        if (rc) return rc;
        for (int k = j - 1; k >= 0; --k) {
            double ang = -PI / pow2f(j - k);
            rc = orc_c_apply(ORC_P, &ang, re, im, n, targets[j], targets[k]);
            if (rc) return rc;
        }
    }
    return ORC_OK;
}

/* ---- reductions ---------------------------------------------------------------------------- */
/* prob0: measurement.rs:16-29 == core.rs:202-215.  Sum over chunks of 2^(t+1) of the first 2^t
   |amp|^2 (powi(2) == x*x).  rayon's reduction order is nondeterministic; this uses per-chunk
   partial sums combined in chunk order. */
ORC_API double orc_prob0(const double *re, const double *im, int n, int target) {
    const idx_t len = (idx_t)1 << n, dist = (idx_t)1 << target, chunk = dist << 1;
    const idx_t nchunks = len / chunk;
    double total = 0.0;
    if (nchunks >= 64) {
        #pragma omp parallel for schedule(static) reduction(+ : total) if (!serial_path(n))
        for (idx_t c = 0; c < nchunks; ++c) {
            double s = 0.0;
            const double *r = re + c * chunk, *m = im + c * chunk;
            for (idx_t j = 0; j < dist; ++j) s += r[j] * r[j] + m[j] * m[j];
            total += s;
        }
    } else {
        for (idx_t c = 0; c < nchunks; ++c) {
            const double *r = re + c * chunk, *m = im + c * chunk;
            double s = 0.0;
            #pragma omp parallel for schedule(static) reduction(+ : s) if (!serial_path(n))
            for (idx_t j = 0; j < dist; ++j) s += r[j] * r[j] + m[j] * m[j];
            total += s;
        }
    }
    return total;
}

ORC_API double orc_norm2(const double *re, const double *im, int n) {
    const idx_t len = (idx_t)1 << n;
    double s = 0.0;
    #pragma omp parallel for schedule(static) reduction(+ : s) if (!serial_path(n))
    for (idx_t i = 0; i < len; ++i) s += re[i] * re[i] + im[i] * im[i];
    return s;
}

/* qubit_expectation_value core.rs:198-219 */
ORC_API double orc_qubit_expectation_value(const double *re, const double *im, int n, int target) {
    return 2.0 * orc_prob0(re, im, n, target) - 1.0;
}

/* measure_qubit measurement.rs:12-92.  forced_v in {0,1} mirrors `v: Some(_)`; forced_v < 0 mirrors
   `None`, with the Bernoulli(1-prob0) draw (measurement.rs:35-36, unseeded thread_rng in the reference)
   replaced by the caller-supplied uniform u01: outcome 1 iff u01 < 1 - prob0. */
ORC_API int orc_measure_qubit(double *re, double *im, int n, int target, int reset, int forced_v,
                              double u01, double *prob0_out) {
    const idx_t len = (idx_t)1 << n, dist = (idx_t)1 << target, chunk = dist << 1;
    double prob0 = orc_prob0(re, im, n, target);
    if (prob0_out) *prob0_out = prob0;
    int val = forced_v >= 0 ? forced_v : (u01 < 1.0 - prob0 ? 1 : 0);
    const idx_t nchunks = len / chunk;
    if (val == 0) {
        const double k = 1.0 / sqrt(prob0); /* prob0.sqrt().recip() measurement.rs:40 */
        #pragma omp parallel for schedule(static) if (!serial_path(n) && nchunks >= 64)
        for (idx_t c = 0; c < nchunks; ++c) {
            double *r = re + c * chunk, *m = im + c * chunk;
            for (idx_t j = 0; j < dist; ++j) {
                r[j] *= k; m[j] *= k; r[j + dist] = 0.0; m[j + dist] = 0.0;
            }
        }
    } else {
        const double prob1 = 1.0 - prob0;
        const double k = 1.0 / sqrt(prob1); /* measurement.rs:63-64 */
        #pragma omp parallel for schedule(static) if (!serial_path(n) && nchunks >= 64)
        for (idx_t c = 0; c < nchunks; ++c) {
            double *r = re + c * chunk, *m = im + c * chunk;
            for (idx_t j = 0; j < dist; ++j) {
                r[j + dist] *= k; m[j + dist] *= k; r[j] = 0.0; m[j] = 0.0;
            }
        }
        if (reset) orc_apply(ORC_X, NULL, 0, 0, re, im, n, target); /* measurement.rs:87-89 */
    }
    return val;
}

/* xyz_expectation_value core.rs:222-264: per target, clone, apply O, Re<psi|O psi> = sum(a*c + b*d).
   Returns ORC_ERR_INVALID for an observable outside "xyz" (the reference panics, core.rs:223-225). */
ORC_API int orc_xyz_expectation_value(char observable, const double *re, const double *im, int n,
                                      const int *targets, int k, double *out) {
    if (!(observable == 'x' || observable == 'y' || observable == 'z')) return ORC_ERR_INVALID;
    const idx_t len = (idx_t)1 << n;
    double *wr = (double *)malloc((size_t)len * sizeof(double));
    double *wi = (double *)malloc((size_t)len * sizeof(double));
    if (!wr || !wi) { free(wr); free(wi); return ORC_ERR_INVALID; }
    for (int t = 0; t < k; ++t) {
        memcpy(wr, re, (size_t)len * sizeof(double));
        memcpy(wi, im, (size_t)len * sizeof(double));
        int kind = observable == 'z' ? ORC_Z : observable == 'y' ? ORC_Y : ORC_X;
        int rc = orc_apply(kind, NULL, 0, 0, wr, wi, n, targets[t]);
        if (rc) { free(wr); free(wi); return rc; }
        double s = 0.0;
        #pragma omp parallel for schedule(static) reduction(+ : s) if (!serial_path(n))
        for (idx_t i = 0; i < len; ++i) s += re[i] * wr[i] + im[i] * wi[i];
        out[t] = s;
    }
    free(wr); free(wi);
    return ORC_OK;
}

/* ---- sampling ------------------------------------------------------------------------------- */
/* splitmix64: the shared deterministic generator of this repo's host-side randomness (the reference
   uses unseeded thread_rng everywhere, core.rs:88,105 -- bitwise parity of random streams is
   impossible by construction). */
static inline uint64_t splitmix64(uint64_t *x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double u01_from(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

ORC_API void orc_uniforms(uint64_t seed, int64_t count, double *out) {
    uint64_t s = seed;
    for (int64_t i = 0; i < count; ++i) out[i] = u01_from(splitmix64(&s));
}

/* Reservoir::update/sampling core.rs:81-112 restated with the seeded generator: num_tests uniform
   index draws, weight |amp|^2, every slot replaces itself with probability w_i / w_s. */
ORC_API void orc_reservoir_sampling(const double *re, const double *im, int n, int64_t k,
                                    int64_t num_tests, uint64_t seed, int64_t *entries) {
    const idx_t len = (idx_t)1 << n;
    uint64_t s = seed;
    double w_s = 0.0;
    for (int64_t i = 0; i < k; ++i) entries[i] = 0;
    for (int64_t t = 0; t < num_tests; ++t) {
        idx_t e = (idx_t)(splitmix64(&s) % (uint64_t)len);
        double md = sqrt(fma(re[e], re[e], im[e] * im[e])); /* modulus math.rs:28-30 */
        double w = md * md;                                 /* .powi(2) core.rs:98 */
        w_s += w;
        double delta = w / w_s;
        for (int64_t j = 0; j < k; ++j) {
            double eps = u01_from(splitmix64(&s));
            if (eps < delta) entries[j] = e;
        }
    }
}

/* Exact inverse-CDF sampling (what the engine's spz_sample computes): outcome for uniform u is the
   smallest i with cdf[i] > u * total, cdf = running sum of |amp|^2 in index order. */
ORC_API void orc_sample_cdf(const double *re, const double *im, int n, const double *u, int64_t shots,
                            int64_t *out) {
    const idx_t len = (idx_t)1 << n;
    double *cdf = (double *)malloc((size_t)len * sizeof(double));
    double acc = 0.0;
    for (idx_t i = 0; i < len; ++i) { acc += re[i] * re[i] + im[i] * im[i]; cdf[i] = acc; }
    for (int64_t s = 0; s < shots; ++s) {
        double x = u[s] * acc;
        idx_t lo = 0, hi = len - 1;
        while (lo < hi) { idx_t mid = (lo + hi) >> 1; if (cdf[mid] > x) hi = mid; else lo = mid + 1; }
        out[s] = lo;
    }
    free(cdf);
}

/* ---- gen_random_state  utils.rs:168-201 with the seeded generator --------------------------- */
ORC_API void orc_gen_random_state(double *re, double *im, int n, uint64_t seed) {
    const double PI = 3.14159265358979323846;
    const idx_t len = (idx_t)1 << n;
    uint64_t s = seed;
    double total = 0.0;
    for (idx_t i = 0; i < len; ++i) { re[i] = u01_from(splitmix64(&s)); total += re[i]; }
    double total_recip = 1.0 / total;
    for (idx_t i = 0; i < len; ++i) {
        double p = re[i] * total_recip;
        double a = u01_from(splitmix64(&s)) * (2.0 * PI);
        double ps = sqrt(p);
        re[i] = ps * cos(a);
        im[i] = ps * sin(a);
    }
}

/* ---- QuantumCircuit::execute  circuit.rs:552-600 --------------------------------------------- */
/* Same layout as spz_op in include/spinoza_b200.h. */
typedef struct {
    int32_t kind;        /* gate kind */
    int32_t target;
    int32_t t0, t1;      /* SWAP operands (gates.rs:67) */
    double p[3];         /* gate parameters */
    int32_t ctrl_kind;   /* 0 None, 1 Single, 2 Ones, 3 Mixed  (circuit.rs:55-70) */
    int32_t reserved;
    uint64_t ctrl_mask;  /* set of control qubits */
    uint64_t zeros_mask; /* Mixed: set of zeros (circuit.rs:68) */
} orc_op;

/* measured/vals mirror QubitTracker (circuit.rs:122-164).  u01[] supplies one uniform per executed
   Gate::M / Gate::BitFlipNoise in program order (the reference draws from thread_rng). */
ORC_API int orc_execute(double *re, double *im, int n, const orc_op *ops, int nops, uint64_t *measured,
                        uint64_t *vals, const double *u01, int nu) {
    int ui = 0;
    for (int i = 0; i < nops; ++i) {
        const orc_op *op = &ops[i];
        int rc = ORC_OK;
        if (op->kind == ORC_UNITARY) return ORC_ERR_UNSUPPORTED; /* transform_u: out of scope */
        if (op->ctrl_kind == 0 && op->kind == ORC_M) { /* circuit.rs:559-566 */
            if (!((*measured >> op->target) & 1)) {
                double u = ui < nu ? u01[ui] : 0.5; ++ui;
                int v = orc_measure_qubit(re, im, n, op->target, 1, -1, u, NULL);
                *measured |= (uint64_t)1 << op->target;
                *vals &= ~((uint64_t)1 << op->target);
                *vals |= (uint64_t)v << op->target;
            }
        } else if (op->ctrl_kind == 0) { /* circuit.rs:567-569 */
            if (op->kind == ORC_BITFLIP) { /* gates.rs:1365-1374: epsilon <= prob -> X */
                double u = ui < nu ? u01[ui] : 0.5; ++ui;
                if (u <= op->p[0]) rc = orc_apply(ORC_X, NULL, 0, 0, re, im, n, op->target);
            } else {
                rc = orc_apply(op->kind, op->p, op->t0, op->t1, re, im, n, op->target);
            }
        } else if (op->ctrl_kind == 1) { /* circuit.rs:570-578 */
            int c = __builtin_ctzll(op->ctrl_mask);
            if ((*measured >> c) & 1) {
                if ((*vals >> c) & 1) rc = orc_apply(op->kind, op->p, op->t0, op->t1, re, im, n, op->target);
            } else {
                rc = orc_c_apply(op->kind, op->p, re, im, n, c, op->target);
            }
        } else if (op->ctrl_kind == 2 && op->kind == ORC_X) {
            /* circuit.rs:579-587 calls cc_apply(controls[0], controls[1]); with exactly two controls
               (the only shape `ccx` builds, circuit.rs:277-283) that is this mask. */
            rc = orc_mc_apply_mask(ORC_X, NULL, re, im, n, op->ctrl_mask, op->target);
        } else if (op->ctrl_kind == 3) { /* circuit.rs:588-596 -> mc_apply, zeros dropped (B4) */
            rc = orc_mc_apply_mask(op->kind, op->p, re, im, n, op->ctrl_mask & ~op->zeros_mask, op->target);
        } else {
            return ORC_ERR_UNSUPPORTED; /* todo!() circuit.rs:597 */
        }
        if (rc) return rc;
    }
    return ORC_OK;
}
