/*
 * fmm_replay_model.c -- CPU model of the DEVICE band-rebuild algorithm (lsml_b200/csrc/fmm.cu).
 * TEST INFRASTRUCTURE ONLY (see oracle/lsml_oracle.c for the rules).
 *
 * fmm.cu does not run a heap: it iterates the local "replay" operator V to its fixed point.
 * This file states the same operator in plain C so that the claim "the fixed point of V is the
 * sequential fast-marching result (orc_fmm_distance)" can be checked on the CPU, on many random
 * inputs, without a GPU (tests/test_fmm_algorithm.py).  It is NOT the oracle for the reference
 * (that is orc_fmm_distance); it is a model of our own kernel.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char u8;

/* Ordering decisions of the marcher (heap order, upwind admissibility, choice of the nearer
 * side, second-order switch) are made on |d| rounded to float32: values that are equal in
 * exact arithmetic but differ by float64 rounding noise (symmetric inputs such as the +-1
 * step function of a ball initializer) then tie exactly and are ordered by flat index, not by
 * noise.  The distances themselves stay float64. */
static inline float qmag(double d) { return (float)fabs(d); }

/* Contributor rule: a second neighbour of OPPOSITE sign lies across the
 * interface -- its magnitude says nothing about causality -- and is exempt from the causality
 * test and from eviction.  With it the marcher's freeze values are monotone in freeze order, which
 * is what makes the sequential result the unique fixed point of the local replay operator the
 * GPU iterates.  LSML_ORACLE_RULE_R2=0 restores the rule used before (largest magnitude evicted
 * regardless of side; fmm.cu: -DLSML_FMM_RULE_LEGACY), kept for the regression test of the
 * difference. */
static int across(double v1, double v2) {
    static int rule = -1;
    if (rule < 0) { const char *e = getenv("LSML_ORACLE_RULE_R2"); rule = (e && e[0] == '0') ? 0 : 1; }
    return rule && v2 * v1 < 0;
}

typedef struct {
    int ndim, shape[3];
    long st[3], size;
    double dx[3], idx2[3];
    const double *phi;
    double *T;
    u8 *state;   /* 0 far, 1 initial frozen, 2 candidate */
    double band;
} rm_t;

static int key_less(double ta, long ia, double tb, long ib) {
    float a = qmag(ta), b = qmag(tb);
    return (a < b) || (a == b && ia < ib);
}

/* same arithmetic as lsml_oracle.c:fmm_solve */
static double rm_solve(int ndim, const double *idx2, const double *v1in, const double *v2in,
                        double phi) {
    const double aa = 9.0 / 4.0, one_third = 1.0 / 3.0;
    double v1[3], v2[3];
    int active[3];
    int nactive = 0;
    for (int ax = 0; ax < ndim; ++ax) {
        v1[ax] = v1in[ax]; v2[ax] = v2in[ax];
        active[ax] = v1[ax] < DBL_MAX; nactive += active[ax];
    }
    while (nactive > 0) {
        double a = 0, b = 0, c = 0;
        for (int ax = 0; ax < ndim; ++ax) {
            if (!active[ax]) continue;
            if (v2[ax] < DBL_MAX) {
                double tp = one_third * (4 * v1[ax] - v2[ax]);
                a += idx2[ax] * aa;
                b -= idx2[ax] * 2 * aa * tp;
                c += idx2[ax] * aa * (tp * tp);
            } else {
                a += idx2[ax];
                b -= idx2[ax] * 2 * v1[ax];
                c += idx2[ax] * (v1[ax] * v1[ax]);
            }
        }
        c -= 1;
        double det = b * b - 4 * a * c;
        if (det > 0) {
            double r = (phi > DBL_EPSILON) ? (-b + sqrt(det)) / 2.0 / a
                                           : (-b - sqrt(det)) / 2.0 / a;
            int causal = 1;
            for (int ax = 0; ax < ndim; ++ax) {
                if (!active[ax]) continue;
                if (qmag(r) <= qmag(v1[ax])) causal = 0;
                if (v2[ax] < DBL_MAX && !across(v1[ax], v2[ax]) && qmag(r) <= qmag(v2[ax])) causal = 0;
            }
            if (causal) return r;
        }
        /* remove the contributor farthest from the interface: a value2 demotes its axis to
         * first order, a value1 drops the axis (first maximum in the order ax0.v2, ax0.v1, ...) */
        int worst = -1, worst_is_v2 = 0;
        float wv = -1.0f;
        for (int ax = 0; ax < ndim; ++ax) {
            if (!active[ax]) continue;
            if (v2[ax] < DBL_MAX && !across(v1[ax], v2[ax]) && qmag(v2[ax]) > wv) { wv = qmag(v2[ax]); worst = ax; worst_is_v2 = 1; }
            if (qmag(v1[ax]) > wv) { wv = qmag(v1[ax]); worst = ax; worst_is_v2 = 0; }
        }
        if (worst_is_v2) v2[worst] = DBL_MAX;
        else { active[worst] = 0; nactive--; }
    }
    return 0.0;
}

static double stencil_update(const rm_t *m, double t1[3][2], double t2[3][2], int f1[3][2],
                             int f2[3][2], double phi) {
    double v1[3], v2[3];
    for (int ax = 0; ax < m->ndim; ++ax) {
        v1[ax] = DBL_MAX; v2[ax] = DBL_MAX;
        for (int d = 0; d < 2; ++d)
            if (f1[ax][d] && (v1[ax] == DBL_MAX || qmag(t1[ax][d]) < qmag(v1[ax]))) {
                v1[ax] = t1[ax][d];
                v2[ax] = DBL_MAX;
                if (f2[ax][d] && (v1[ax] == 0.0 || t2[ax][d] * v1[ax] < 0 ||
                                  qmag(t2[ax][d]) <= qmag(v1[ax])))
                    v2[ax] = t2[ax][d];
            }
    }
    return rm_solve(m->ndim, m->idx2, v1, v2, phi);
}

static double replay(const rm_t *m, long l) {
    double t1[3][2], t2[3][2];
    int f1[3][2], f2[3][2], e1[3][2], e2[3][2];
    int any = 0;
    for (int ax = 0; ax < m->ndim; ++ax) {
        long c = (l / m->st[ax]) % m->shape[ax];
        for (int d = 0; d < 2; ++d) {
            int j = d ? 1 : -1;
            t1[ax][d] = t2[ax][d] = INFINITY;
            f1[ax][d] = f2[ax][d] = e1[ax][d] = e2[ax][d] = 0;
            if (c + j >= 0 && c + j < m->shape[ax]) {
                long n = l + j * m->st[ax];
                t1[ax][d] = m->T[n];
                f1[ax][d] = m->state[n] == 1;
                e1[ax][d] = m->state[n] == 2 && fabs(m->T[n]) <= m->band;
                any |= f1[ax][d];
            }
            if (c + 2 * j >= 0 && c + 2 * j < m->shape[ax]) {
                long n = l + 2 * j * m->st[ax];
                t2[ax][d] = m->T[n];
                f2[ax][d] = m->state[n] == 1;
                e2[ax][d] = m->state[n] == 2 && fabs(m->T[n]) <= m->band;
            }
        }
    }
    double phi = m->phi[l], t = INFINITY;
    int narrow = 0;
    if (any) {
        double d0 = stencil_update(m, t1, t2, f1, f2, phi);
        if (d0 != 0.0) { t = d0; narrow = 1; }
    }
    for (int it = 0; it < 4 * m->ndim; ++it) {
        int bax = -1, bd = 0, bw = 0; double bt = INFINITY; long bi = 0;
        for (int ax = 0; ax < m->ndim; ++ax)
            for (int d = 0; d < 2; ++d) {
                int j = d ? 1 : -1;
                if (e1[ax][d]) {
                    long n = l + j * m->st[ax];
                    if (bax < 0 || key_less(t1[ax][d], n, bt, bi)) { bax = ax; bd = d; bw = 1; bt = t1[ax][d]; bi = n; }
                }
                if (e2[ax][d]) {
                    long n = l + 2 * j * m->st[ax];
                    if (bax < 0 || key_less(t2[ax][d], n, bt, bi)) { bax = ax; bd = d; bw = 2; bt = t2[ax][d]; bi = n; }
                }
            }
        if (bax < 0) break;
        if (narrow && key_less(t, l, bt, bi)) break;
        int reeval;
        if (bw == 1) { e1[bax][bd] = 0; f1[bax][bd] = 1; reeval = 1; }
        else { e2[bax][bd] = 0; f2[bax][bd] = 1; reeval = f1[bax][bd] && narrow; }
        if (reeval) {
            double dn = stencil_update(m, t1, t2, f1, f2, phi);
            if (dn != 0.0) { t = dn; narrow = 1; }
        }
    }
    return narrow ? t : INFINITY;
}

/* jacobi != 0: synchronous sweeps (all voxels read the previous sweep's values);
 * jacobi == 0: in-place sweeps in list order.  Returns sweeps used, or -sweeps if the cap was
 * hit.  changed_out (size = total, may be NULL) flags voxels still changing in the last sweep. */
long rm_fmm_fixed_point(int ndim, const int *shape, const double *phi, const double *dx,
                        double narrow, int jacobi, int max_sweeps, double *dist_out,
                        u8 *frozen_out, u8 *changed_out) {
    rm_t m;
    memset(&m, 0, sizeof m);
    m.ndim = ndim; m.size = 1;
    for (int a = 0; a < ndim; ++a) { m.shape[a] = shape[a]; m.size *= shape[a]; }
    long s = 1;
    for (int a = ndim - 1; a >= 0; --a) { m.st[a] = s; s *= shape[a]; }
    for (int a = 0; a < ndim; ++a) { m.dx[a] = dx[a]; m.idx2[a] = 1.0 / dx[a] / dx[a]; }
    m.phi = phi;
    m.band = narrow != 0 ? narrow : DBL_MAX;
    m.T = (double *)malloc(sizeof(double) * m.size);
    m.state = (u8 *)calloc(m.size, 1);
    double *Tn = (double *)malloc(sizeof(double) * m.size);
    long *list = (long *)malloc(sizeof(long) * m.size), nc = 0;
    for (long i = 0; i < m.size; ++i) m.T[i] = INFINITY;
    /* initial frozen set */
    for (long i = 0; i < m.size; ++i) {
        if (phi[i] == 0.0) { m.T[i] = 0.0; m.state[i] = 1; continue; }
        int borders = 0; double dsum = 0;
        for (int ax = 0; ax < ndim; ++ax) {
            long c = (i / m.st[ax]) % m.shape[ax];
            double ld = 0;
            for (int j = -1; j < 2; j += 2) {
                if (c + j < 0 || c + j >= m.shape[ax]) continue;
                double nv = phi[i + j * m.st[ax]];
                if (phi[i] * nv < 0) {
                    borders = 1;
                    double d = dx[ax] * phi[i] / (phi[i] - nv);
                    if (ld == 0 || ld > d) ld = d;
                }
            }
            if (ld > 0) dsum += 1 / ld / ld;
        }
        if (borders) { m.T[i] = phi[i] < 0 ? -sqrt(1 / dsum) : sqrt(1 / dsum); m.state[i] = 1; }
    }
#define PUSH_NEIGHBOURS(l) \
    for (int ax = 0; ax < ndim; ++ax) { \
        long c = ((l) / m.st[ax]) % m.shape[ax]; \
        for (int j = -1; j < 2; j += 2) { \
            if (c + j < 0 || c + j >= m.shape[ax]) continue; \
            long n = (l) + j * m.st[ax]; \
            if (m.state[n] == 0) { m.state[n] = 2; list[nc++] = n; } \
        } }
    for (long i = 0; i < m.size; ++i) if (m.state[i] == 1) { PUSH_NEIGHBOURS(i) }
    long sweep = 0; int conv = 0;
    for (; sweep < max_sweeps; ++sweep) {
        long n0 = nc; int changed = 0;
        if (changed_out) memset(changed_out, 0, m.size);
        for (long t = 0; t < n0; ++t) {
            long l = list[t];
            double old = m.T[l], nw = replay(&m, l);
            if (jacobi) Tn[l] = nw; else m.T[l] = nw;
            if (memcmp(&old, &nw, sizeof old) != 0) { changed = 1; if (changed_out) changed_out[l] = 1; }
            if (fabs(nw) <= m.band && !(fabs(old) <= m.band)) { PUSH_NEIGHBOURS(l) }
        }
        if (jacobi) for (long t = 0; t < n0; ++t) m.T[list[t]] = Tn[list[t]];
        if (!changed && nc == n0) { conv = 1; ++sweep; break; }
    }
    for (long i = 0; i < m.size; ++i) {
        int fr = m.state[i] == 1 || (m.state[i] == 2 && fabs(m.T[i]) <= m.band);
        frozen_out[i] = (u8)fr;
        dist_out[i] = fr ? m.T[i] : 0.0;
    }
    free(m.T); free(m.state); free(Tn); free(list);
    return conv ? sweep : -sweep;
}
