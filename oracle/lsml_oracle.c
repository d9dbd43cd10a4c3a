/*
 * lsml_oracle.c -- CPU restatement of the lsml level-set hot path (TEST INFRASTRUCTURE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (lsml_b200) never does.
 *
 * What is restated, and from where (paths relative to /root/reference):
 *
 *   orc_gmag_os            lsml/util/_cutil/masked_gradient.c:152-315  (gmag_os{3,2,1}d)
 *   orc_gradient_centered  lsml/util/_cutil/masked_gradient.c:13-150   (gradient_centered{3,2,1}d)
 *   orc_fmm_distance       scikit-fmm `skfmm.distance(phi, dx, narrow)` as called by
 *                          lsml/util/distance_transform.py:49 -- a THIRD-PARTY dependency
 *                          (PyPI scikit_fmm, unpinned in setup.py:59) whose source is NOT
 *                          under /root/reference.  PARITY UNPINNED beyond the reference's own
 *                          known-answer tests (lsml/util/test/test_distance_transform.py:11-53):
 *                          the algorithm below is the published fast-marching method with the
 *                          second-order upwind stencil as scikit-fmm implements it, written
 *                          from the description in SURVEY.md section 8c.  Two points where the
 *                          upstream behaviour is not known for certain, or is not well defined,
 *                          are decided here and documented at the code: heap ties (broken by
 *                          flat index) and non-causal quadratic roots (the farthest neighbour
 *                          is dropped and the quadratic re-solved, as the published method
 *                          prescribes; see fmm_solve).
 *   orc_marching_cubes_area
 *                          surface area of {u = 0} in 3-D (skimage marching_cubes_lewiner +
 *                          mesh_surface_area, lsml/feature/provided/shape.py:88-89); see the
 *                          comment at the function: PARITY UNPINNED beyond test_shape.py:59-107.
 *   orc_marching_squares_length
 *                          skimage.measure.find_contours(u, 0) total poly-line length with
 *                          lsml's closing chord and dx[::-1] scaling, lsml/feature/provided/shape.py:70-85.
 *                          Third-party (scikit-image, unpinned, absent): PARITY UNPINNED beyond
 *                          lsml/feature/provided/test/test_shape.py:43-57.
 *
 * The arithmetic is written so that every floating-point operation happens in the same order
 * as in the reference (no re-association); compile with -ffp-contract=off.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char u8;

/* Ordering decisions of the marcher (heap order, upwind admissibility, choice of the nearer
 * side, second-order switch) are made on |d| rounded to float32: values that are equal in
 * exact arithmetic but differ by float64 rounding noise (symmetric inputs such as the +-1
 * step function of a ball initializer) then tie exactly and are ordered by flat index, not by
 * noise.  The distances themselves stay float64. */
static inline float qmag(double d) { return (float)fabs(d); }

/* Ablation switches (tests/tools/band_gap_report.py --ablate): each bit turns ONE of the four
 * deliberate departures from upstream scikit-fmm back into upstream's behaviour, so that the gap
 * to oracle/skfmm_literal.c can be attributed.  0 (the default) is the marcher the product
 * matches; nothing but the report ever sets it. */
enum { ABL_KEY64 = 1, ABL_NO_CAUSALITY = 2, ABL_KEEP_V2 = 4, ABL_BREAK = 8 };
static int g_ablate = 0;
void orc_fmm_set_ablation(int flags) { g_ablate = flags; }
/* ordering magnitude: float32-rounded (default) or the float64 value itself */
static inline double omag(double d) { return (g_ablate & ABL_KEY64) ? fabs(d) : (double)qmag(d); }

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

static inline double dmax(double a, double b) { return a < b ? b : a; } /* helpers.c:57 */
static inline double dmin(double a, double b) { return a > b ? b : a; } /* helpers.c:58 */
static inline double sq(double a) { return a * a; }

/* row-major strides (in elements) of an ndim<=3 grid */
static void strides_of(int ndim, const int *shape, long *st) {
    long s = 1;
    for (int a = ndim - 1; a >= 0; --a) { st[a] = s; s *= shape[a]; }
}

/* ------------------------------------------------------------------------------------------
 * Osher-Sethian upwind gradient magnitude, masked.  One generic routine replaces the three
 * hand-unrolled reference functions; the per-voxel arithmetic (difference, divide by the
 * spacing, select on nu<0, left-to-right sum of squares in axis order b,f) is identical.
 * masked_gradient.c:165-222.
 * ---------------------------------------------------------------------------------------- */
void orc_gmag_os(int ndim, const int *shape, const double *A, const u8 *mask,
                 const double *nu, double *gmag, const double *dx) {
    long st[3], total = 1;
    strides_of(ndim, shape, st);
    for (int a = 0; a < ndim; ++a) total *= shape[a];
    for (long l = 0; l < total; ++l) {
        if (!mask[l]) continue;
        long rem = l;
        double f[3], b[3];
        for (int a = 0; a < ndim; ++a) {
            long c = rem / st[a];
            rem -= c * st[a];
            if (c == 0) {
                f[a] = A[l + st[a]] - A[l];
                b[a] = f[a];
            } else if (c == shape[a] - 1) {
                b[a] = A[l] - A[l - st[a]];
                f[a] = b[a];
            } else {
                f[a] = A[l + st[a]] - A[l];
                b[a] = A[l] - A[l - st[a]];
            }
            f[a] = f[a] / dx[a];
            b[a] = b[a] / dx[a];
        }
        /* left-to-right sum in the reference's order: b-term then f-term, axis by axis */
        double s = 0.0;
        for (int a = 0; a < ndim; ++a) {
            double tb, tf;
            if (nu[l] < 0) { tb = sq(dmax(b[a], 0)); tf = sq(dmin(f[a], 0)); }
            else           { tb = sq(dmin(b[a], 0)); tf = sq(dmax(f[a], 0)); }
            s = (a == 0) ? (tb + tf) : ((s + tb) + tf);
        }
        gmag[l] = sqrt(s);
    }
}

/* ------------------------------------------------------------------------------------------
 * Centered differences, masked.  grads = ndim arrays laid out back to back (grads + a*total).
 * masked_gradient.c:24-71 (3d), 89-119 (2d), 132-148 (1d; magnitude is |d|).
 * ---------------------------------------------------------------------------------------- */
void orc_gradient_centered(int ndim, const int *shape, const double *A, const u8 *mask,
                           double *grads, double *gmag, const double *dx, int normalize) {
    long st[3], total = 1;
    strides_of(ndim, shape, st);
    for (int a = 0; a < ndim; ++a) total *= shape[a];
    for (long l = 0; l < total; ++l) {
        if (!mask[l]) continue;
        long rem = l;
        double d[3];
        for (int a = 0; a < ndim; ++a) {
            long c = rem / st[a];
            rem -= c * st[a];
            if (c == 0) d[a] = A[l + st[a]] - A[l];
            else if (c == shape[a] - 1) d[a] = A[l] - A[l - st[a]];
            else d[a] = 0.5 * (A[l + st[a]] - A[l - st[a]]);
        }
        for (int a = 0; a < ndim; ++a) d[a] = d[a] / dx[a];
        double g;
        if (ndim == 1) {
            g = (d[0] > 0) ? d[0] : -d[0];
        } else {
            double s = sq(d[0]) + sq(d[1]);
            if (ndim == 3) s = s + sq(d[2]);
            g = sqrt(s);
        }
        gmag[l] = g;
        if (normalize == 1 && g > 0)
            for (int a = 0; a < ndim; ++a) d[a] /= g;
        for (int a = 0; a < ndim; ++a) grads[(long)a * total + l] = d[a];
    }
}

/* ==========================================================================================
 * Narrow-band signed distance by the fast-marching method (scikit-fmm `distance`).
 * ======================================================================================== */
enum { FAR = 0, NARROW = 1, FROZEN = 2 };

typedef struct {
    int ndim;
    int shape[3];
    long st[3];
    long size;
    double idx2[3];
    const double *phi;
    double *dist;
    signed char *flag;
    /* binary min-heap on (|dist|, flat index); pos[] = heap slot of a voxel or -1 */
    long *heap;
    long *pos;
    long nheap;
} fmm_t;

/* flat index of the voxel `step` cells along `axis` from i, or -1 when outside the grid */
static long nb(const fmm_t *m, long i, int axis, int step) {
    long c = (i / m->st[axis]) % m->shape[axis];
    long cc = c + step;
    if (cc < 0 || cc >= m->shape[axis]) return -1;
    return i + (long)step * m->st[axis];
}

/* strict total order used by the heap: smaller |d| first, ties by smaller flat index.
 * (scikit-fmm's own tie behaviour depends on its heap internals and is not reproducible
 * without its source; a total order makes the result well defined.) */
static int key_less(const fmm_t *m, long i, long j) {
    double a = omag(m->dist[i]), b = omag(m->dist[j]);
    return (a < b) || (a == b && i < j);
}
static void heap_swap(fmm_t *m, long a, long b) {
    long t = m->heap[a]; m->heap[a] = m->heap[b]; m->heap[b] = t;
    m->pos[m->heap[a]] = a; m->pos[m->heap[b]] = b;
}
static void sift_up(fmm_t *m, long s) {
    while (s > 0) {
        long p = (s - 1) / 2;
        if (key_less(m, m->heap[s], m->heap[p])) { heap_swap(m, s, p); s = p; } else break;
    }
}
static void sift_down(fmm_t *m, long s) {
    for (;;) {
        long l = 2 * s + 1, r = l + 1, b = s;
        if (l < m->nheap && key_less(m, m->heap[l], m->heap[b])) b = l;
        if (r < m->nheap && key_less(m, m->heap[r], m->heap[b])) b = r;
        if (b == s) break;
        heap_swap(m, s, b); s = b;
    }
}
static void heap_push(fmm_t *m, long i) {
    m->heap[m->nheap] = i; m->pos[i] = m->nheap; m->nheap++;
    sift_up(m, m->pos[i]);
}
static void heap_fix(fmm_t *m, long i) { sift_up(m, m->pos[i]); sift_down(m, m->pos[i]); }
static long heap_pop(fmm_t *m) {
    long top = m->heap[0];
    m->nheap--;
    if (m->nheap > 0) { m->heap[0] = m->heap[m->nheap]; m->pos[m->heap[0]] = 0; sift_down(m, 0); }
    m->pos[top] = -1;
    return top;
}

/* Upwind update of voxel i from its currently FROZEN stencil (scikit-fmm
 * distanceMarcher::updatePointOrderTwo + solveQuadratic).  Per axis the frozen neighbour with
 * the smaller |d| supplies value1; when the voxel two cells away in the same direction is
 * frozen too and not farther from the interface (in the signed sense) the second-order
 * one-sided difference (3T - 4 v1 + v2)/(2 dx) is used, i.e. the term
 * 9/4 (T - (4 v1 - v2)/3)^2 / dx^2, else the first-order term (T - v1)^2 / dx^2.
 *
 * CAUSALITY (decided here, see the file header): the published fast-marching method is an
 * upwind scheme -- a value may only be computed from neighbours that are not farther from the
 * interface than the value itself.  When the quadratic has no positive discriminant, or its
 * root is smaller in magnitude than one of the values it was computed from, the contributor
 * with the largest magnitude is removed (a value2: the axis falls back to first order; a
 * value1: the axis is dropped) and the quadratic is solved again (Sethian 1996, the usual
 * implementation of max(D-T, -D+T, 0)).  Upstream scikit-fmm does not make this check; its
 * result then depends on the order in which its heap happens to freeze voxels, which no
 * specification pins down.  A value2 of OPPOSITE sign (an initial frozen voxel across the
 * interface) takes part in neither the test nor the removal -- see across().  With the check
 * every freeze value is >= the values frozen before it, the marcher's result is independent of
 * heap internals and a parallel solver can reproduce it exactly.  Returns 0.0 when there is no
 * frozen neighbour at all. */
static double fmm_solve(int ndim, const double *idx2, const double *v1in, const double *v2in,
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
                if (omag(r) <= omag(v1[ax])) causal = 0;
                if (v2[ax] < DBL_MAX && !across(v1[ax], v2[ax]) && omag(r) <= omag(v2[ax])) causal = 0;
            }
            if (causal || (g_ablate & ABL_NO_CAUSALITY)) return r;
        }
        if (g_ablate & ABL_NO_CAUSALITY) return -b / (2.0 * a);   /* upstream's det < 0 fallback */
        /* remove the contributor farthest from the interface: a value2 demotes its axis to
         * first order, a value1 drops the axis (first maximum in the order ax0.v2, ax0.v1, ...) */
        int worst = -1, worst_is_v2 = 0;
        double wv = -1.0;
        for (int ax = 0; ax < ndim; ++ax) {
            if (!active[ax]) continue;
            if (v2[ax] < DBL_MAX && !across(v1[ax], v2[ax]) && omag(v2[ax]) > wv) { wv = omag(v2[ax]); worst = ax; worst_is_v2 = 1; }
            if (omag(v1[ax]) > wv) { wv = omag(v1[ax]); worst = ax; worst_is_v2 = 0; }
        }
        if (worst_is_v2) v2[worst] = DBL_MAX;
        else { active[worst] = 0; nactive--; }
    }
    return 0.0;
}

/* scikit-fmm's switch `(d2 <= v1 && v1 >= 0) || (d2 >= v1 && v1 <= 0)`: the voxel two cells
 * away must not be farther from the interface than the direct neighbour, in the signed sense
 * (a voxel on the other side of the interface always qualifies). */
static int second_order_ok(double d2, double v1) {
    return v1 == 0.0 || d2 * v1 < 0 || omag(d2) <= omag(v1);
}

static double fmm_update(const fmm_t *m, long i) {
    double v1[3], v2[3];
    for (int ax = 0; ax < m->ndim; ++ax) {
        v1[ax] = DBL_MAX; v2[ax] = DBL_MAX;
        for (int j = -1; j < 2; j += 2) {
            long n = nb(m, i, ax, j);
            if (n != -1 && m->flag[n] == FROZEN &&
                (v1[ax] == DBL_MAX || omag(m->dist[n]) < omag(v1[ax]))) {
                v1[ax] = m->dist[n];
                if (!(g_ablate & ABL_KEEP_V2)) v2[ax] = DBL_MAX;
                long n2 = nb(m, i, ax, 2 * j);
                if (n2 != -1 && m->flag[n2] == FROZEN && second_order_ok(m->dist[n2], v1[ax]))
                    v2[ax] = m->dist[n2];
            }
        }
    }
    return fmm_solve(m->ndim, m->idx2, v1, v2, m->phi[i]);
}

/* phi: level-set values; narrow: band half-width (0 => march the whole grid).
 * dist_out: signed distance, 0 where not frozen (skfmm's masked data after
 * pfmm.post_process_result); frozen_out: 1 where the voxel was frozen (= ~dist.mask).
 * Returns the number of frozen voxels, or -1 if there is no zero contour at all. */
long orc_fmm_distance(int ndim, const int *shape, const double *phi, const double *dx,
                      double narrow, double *dist_out, u8 *frozen_out) {
    fmm_t m;
    memset(&m, 0, sizeof m);
    m.ndim = ndim;
    m.size = 1;
    for (int a = 0; a < ndim; ++a) { m.shape[a] = shape[a]; m.size *= shape[a]; }
    strides_of(ndim, shape, m.st);
    for (int a = 0; a < ndim; ++a) m.idx2[a] = 1.0 / dx[a] / dx[a];
    m.phi = phi;
    m.dist = dist_out;
    m.flag = (signed char *)calloc(m.size, 1);
    m.heap = (long *)malloc(sizeof(long) * m.size);
    m.pos = (long *)malloc(sizeof(long) * m.size);
    for (long i = 0; i < m.size; ++i) { m.pos[i] = -1; m.dist[i] = 0.0; }

    /* -- initial frozen set: exact zeros, and every voxel with an axis neighbour of the
     *    opposite sign; d = sign / sqrt(sum_axes 1/d_a^2), d_a the nearer of the two linear
     *    crossings along axis a.  */
    long nfrozen = 0;
    for (long i = 0; i < m.size; ++i)
        if (phi[i] == 0.0) { m.flag[i] = FROZEN; m.dist[i] = 0.0; nfrozen++; }
    for (long i = 0; i < m.size; ++i) {
        if (m.flag[i] != FAR) continue;
        double ld[3] = {0, 0, 0};
        int borders = 0;
        for (int ax = 0; ax < ndim; ++ax)
            for (int j = -1; j < 2; j += 2) {
                long n = nb(&m, i, ax, j);
                if (n != -1 && phi[i] * phi[n] < 0) {
                    borders = 1;
                    double d = dx[ax] * phi[i] / (phi[i] - phi[n]);
                    if (ld[ax] == 0 || ld[ax] > d) ld[ax] = d;
                }
            }
        if (borders) {
            double dsum = 0;
            for (int ax = 0; ax < ndim; ++ax)
                if (ld[ax] > 0) dsum += 1 / ld[ax] / ld[ax];
            m.dist[i] = (phi[i] < 0) ? -sqrt(1 / dsum) : sqrt(1 / dsum);
            m.flag[i] = FROZEN;   /* flagged in place: later voxels of this loop only test phi */
            nfrozen++;
        }
    }
    if (nfrozen == 0) { free(m.flag); free(m.heap); free(m.pos); return -1; }

    /* -- initial narrow band: far voxels with a frozen axis neighbour */
    for (long i = 0; i < m.size; ++i) {
        if (m.flag[i] != FAR) continue;
        int touch = 0;
        for (int ax = 0; ax < ndim && !touch; ++ax)
            for (int j = -1; j < 2; j += 2) {
                long n = nb(&m, i, ax, j);
                if (n != -1 && m.flag[n] == FROZEN) { touch = 1; break; }
            }
        if (!touch) continue;
        double d = fmm_update(&m, i);
        if (d != 0.0) { m.dist[i] = d; m.flag[i] = NARROW; heap_push(&m, i); }
    }

    /* -- march */
    while (m.nheap > 0) {
        long x = heap_pop(&m);
        /* outside the band: never frozen.  scikit-fmm breaks out of its loop here; with heap
         * order decided on rounded keys (qmag) a later pop can tie with this one and still be
         * inside the band, so the loop continues (nothing is pushed from a discarded voxel). */
        if (narrow != 0 && fabs(m.dist[x]) > narrow) {
            m.flag[x] = NARROW;
            if (g_ablate & ABL_BREAK) break;
            continue;
        }
        m.flag[x] = FROZEN;
        nfrozen++;
        for (int ax = 0; ax < ndim; ++ax)
            for (int j = -1; j < 2; j += 2) {
                long n = nb(&m, x, ax, j);
                if (n != -1 && m.flag[n] != FROZEN) {
                    double d = fmm_update(&m, n);
                    if (d != 0.0) {
                        m.dist[n] = d;
                        if (m.flag[n] == NARROW) heap_fix(&m, n);
                        else { m.flag[n] = NARROW; heap_push(&m, n); }
                    }
                }
                /* second-order stencil: the voxel two cells away sees x through a frozen one */
                if (n != -1 && m.flag[n] == FROZEN) {
                    long n2 = nb(&m, x, ax, 2 * j);
                    if (n2 != -1 && m.flag[n2] == NARROW) {
                        double d = fmm_update(&m, n2);
                        if (d != 0.0) { m.dist[n2] = d; heap_fix(&m, n2); }
                    }
                }
            }
    }
    for (long i = 0; i < m.size; ++i) {
        int fr = (m.flag[i] == FROZEN);
        frozen_out[i] = (u8)fr;
        if (!fr) m.dist[i] = 0.0;
    }
    free(m.flag); free(m.heap); free(m.pos);
    return nfrozen;
}

/* ==========================================================================================
 * Marching squares (skimage.measure.find_contours, level 0, fully_connected='low',
 * positive_orientation irrelevant for lengths), summed as lsml's BoundarySize does.
 * Contour vertices are (row, col) and are scaled by dx[::-1] (shape.py:80).  Contours are
 * assembled so that the closing chord of shape.py:79 (last vertex back to the first, zero for
 * closed contours, non-zero for contours that end on the image border) can be added.
 * ======================================================================================== */
static double frac(double from, double to) { /* _get_fraction with level 0 */
    if (to == from) return 0.0;
    return (0.0 - from) / (to - from);
}

typedef struct { double r, c; } pt_t;

/* Edge ids: every crossing point lies on a unique grid edge.  Horizontal edge (r, c)-(r, c+1)
 * has id 2*(r*n+c); vertical edge (r, c)-(r+1, c) has id 2*(r*n+c)+1. */
typedef struct { long e0, e1; pt_t p0, p1; } seg_t;

double orc_marching_squares_length(int m_, int n_, const double *u, const double *dx,
                                   int closing_chord) {
    long m = m_, n = n_;
    long cap = 2 * (m - 1) * (n - 1) + 1, ns = 0;
    seg_t *segs = (seg_t *)malloc(sizeof(seg_t) * (cap > 0 ? cap : 1));
    for (long r0 = 0; r0 + 1 < m; ++r0)
        for (long c0 = 0; c0 + 1 < n; ++c0) {
            long r1 = r0 + 1, c1 = c0 + 1;
            double ul = u[r0 * n + c0], ur = u[r0 * n + c1];
            double ll = u[r1 * n + c0], lr = u[r1 * n + c1];
            if (isnan(ul) || isnan(ur) || isnan(ll) || isnan(lr)) continue;
            int sc = (ul > 0) + 2 * (ur > 0) + 4 * (ll > 0) + 8 * (lr > 0);
            if (sc == 0 || sc == 15) continue;
            pt_t top = {(double)r0, c0 + frac(ul, ur)};
            pt_t bottom = {(double)r1, c0 + frac(ll, lr)};
            pt_t left = {r0 + frac(ul, ll), (double)c0};
            pt_t right = {r0 + frac(ur, lr), (double)c1};
            long et = 2 * (r0 * n + c0), eb = 2 * (r1 * n + c0);
            long el = 2 * (r0 * n + c0) + 1, er = 2 * (r0 * n + c1) + 1;
#define SEG(A, EA, B, EB) do { segs[ns].p0 = A; segs[ns].e0 = EA; segs[ns].p1 = B; segs[ns].e1 = EB; ns++; } while (0)
            switch (sc) {
            case 1: SEG(top, et, left, el); break;
            case 2: SEG(right, er, top, et); break;
            case 3: SEG(right, er, left, el); break;
            case 4: SEG(left, el, bottom, eb); break;
            case 5: SEG(top, et, bottom, eb); break;
            case 6: SEG(right, er, top, et); SEG(left, el, bottom, eb); break; /* 'low' */
            case 7: SEG(right, er, bottom, eb); break;
            case 8: SEG(bottom, eb, right, er); break;
            case 9: SEG(top, et, left, el); SEG(bottom, eb, right, er); break; /* 'low' */
            case 10: SEG(bottom, eb, top, et); break;
            case 11: SEG(bottom, eb, left, el); break;
            case 12: SEG(left, el, right, er); break;
            case 13: SEG(top, et, right, er); break;
            case 14: SEG(left, el, top, et); break;
            }
#undef SEG
        }
    double total = 0.0;
    for (long s = 0; s < ns; ++s) {
        double dr = (segs[s].p1.r - segs[s].p0.r) * dx[1];
        double dc = (segs[s].p1.c - segs[s].p0.c) * dx[0];
        total += sqrt(dr * dr + dc * dc);
    }
    if (closing_chord && ns > 0) {
        /* Open contours start and end on edges used by exactly one segment.  Walk each open
         * chain from one free end to the other and add the chord between the two ends. */
        long nedges = 2 * m * n;
        long *first = (long *)malloc(sizeof(long) * nedges);
        long *second = (long *)malloc(sizeof(long) * nedges);
        for (long e = 0; e < nedges; ++e) first[e] = second[e] = -1;
        for (long s = 0; s < ns; ++s) {
            long es[2] = {segs[s].e0, segs[s].e1};
            for (int k = 0; k < 2; ++k) {
                if (first[es[k]] < 0) first[es[k]] = s; else second[es[k]] = s;
            }
        }
        u8 *seen = (u8 *)calloc(ns, 1);
        for (long s = 0; s < ns; ++s) {
            for (int k = 0; k < 2; ++k) {
                long e = k ? segs[s].e1 : segs[s].e0;
                if (second[e] >= 0 || seen[s]) continue;
                /* s has a free end at e: walk to the other free end */
                pt_t start = k ? segs[s].p1 : segs[s].p0;
                long cur = s, ein = e;
                pt_t end = start;
                for (;;) {
                    seen[cur] = 1;
                    long eout = (segs[cur].e0 == ein) ? segs[cur].e1 : segs[cur].e0;
                    end = (segs[cur].e0 == ein) ? segs[cur].p1 : segs[cur].p0;
                    long nxt = (first[eout] == cur) ? second[eout] : first[eout];
                    if (nxt < 0) break;
                    cur = nxt; ein = eout;
                }
                double dr = (end.r - start.r) * dx[1], dc = (end.c - start.c) * dx[0];
                total += sqrt(dr * dr + dc * dc);
            }
        }
        free(first); free(second); free(seen);
    }
    free(segs);
    return total;
}

/* ==========================================================================================
 * Surface area of the zero level set by marching cubes (BoundarySize ndim=3, shape.py:88-89:
 * skimage.measure.marching_cubes_lewiner(u, 0., spacing=dx) + mesh_surface_area).
 *
 * THIRD PARTY, ABSENT, UNPINNED: scikit-image is not installed here, its Lewiner lookup tables
 * (33 topological cases, ~2000 table rows) are not under /root/reference and cannot be restated
 * from memory.  What is restated is the geometry those tables encode wherever it is unambiguous,
 * plus Lewiner's FACE test for the ambiguous faces; the interior ("tunnel") test of cases 4, 6, 7,
 * 10, 12, 13 and the extra centre vertex some of their sub-cases use are NOT reproduced.  PARITY
 * UNPINNED beyond lsml/feature/provided/test/test_shape.py:59-107 (sphere area 4 pi to 0 places,
 * sphericity 1 to 2 places).  skimage also converts the volume to float32 first; this restatement
 * stays in float64.
 *
 * Specification (the device kernel lsml_b200/csrc/mcubes.cu implements the same one, written
 * independently):
 *   - cube corner (d0,d1,d2) in {0,1}^3 is INSIDE when u > 0;
 *   - edge along axis a at the other two coordinates (dp,dq), p<q, has id 4a + 2 dp + dq; a cut
 *     edge carries the vertex corner_lo + t e_a, t = u_lo / (u_lo - u_hi), scaled by dx;
 *   - on each of the 6 faces the cut edges are paired into segments: two cut edges -> one segment;
 *     four (corners alternate in sign) -> the asymptotic decider: the inside corners are joined
 *     through the face iff (product of the two inside values) > (product of the two outside
 *     values), and then each OUTSIDE corner is cut off by a segment between its two edges;
 *     otherwise each inside corner is cut off;
 *   - every cut edge lies on two faces, so the segments close into loops; each loop is
 *     triangulated as a fan from its vertex with the lowest edge id, walking towards the
 *     neighbour found first (faces in the order axis 0 lo, axis 0 hi, axis 1 lo, ...);
 *   - triangle area = 0.5 * sqrt((cx^2 + cy^2) + cz^2) of the cross product of (p_i - p_0) and
 *     (p_{i+1} - p_0); the areas are added in loop / fan order per cube, cubes in flat order.
 * ======================================================================================== */
static int mc_edge_id(int axis, int dp, int dq) { return 4 * axis + 2 * dp + dq; }

static double mc_cube_area(const double v[2][2][2], const double *dx) {
    int inside[2][2][2], any_in = 0, any_out = 0;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) {
        inside[a][b][c] = v[a][b][c] > 0;
        if (inside[a][b][c]) any_in = 1; else any_out = 1;
    }
    if (!any_in || !any_out) return 0.0;
    double pos[12][3];
    int cut[12];
    for (int e = 0; e < 12; ++e) cut[e] = 0;
    for (int axis = 0; axis < 3; ++axis)
        for (int dp = 0; dp < 2; ++dp) for (int dq = 0; dq < 2; ++dq) {
            int lo[3], hi[3], p = axis == 0 ? 1 : 0, q = axis == 2 ? 1 : 2;
            lo[axis] = 0; hi[axis] = 1; lo[p] = hi[p] = dp; lo[q] = hi[q] = dq;
            double vl = v[lo[0]][lo[1]][lo[2]], vh = v[hi[0]][hi[1]][hi[2]];
            int e = mc_edge_id(axis, dp, dq);
            if ((vl > 0) != (vh > 0)) {
                cut[e] = 1;
                double t = vl / (vl - vh);
                for (int k = 0; k < 3; ++k) pos[e][k] = (double)lo[k] * dx[k];
                pos[e][axis] = t * dx[axis];
            }
        }
    int nb[12][2], nnb[12];
    for (int e = 0; e < 12; ++e) nnb[e] = 0;
#define MC_LINK(E, F) do { nb[E][nnb[E]++] = (F); nb[F][nnb[F]++] = (E); } while (0)
    for (int axis = 0; axis < 3; ++axis)
        for (int s = 0; s < 2; ++s) {
            int p = axis == 0 ? 1 : 0, q = axis == 2 ? 1 : 2;
            /* corners of the face in cyclic order (dp,dq) = 00, 01, 11, 10 and the face edge
             * BETWEEN corner i and corner i+1 */
            const int cp[4] = {0, 0, 1, 1}, cq[4] = {0, 1, 1, 0};
            int fe[4], in4[4];
            double val[4];
            for (int i = 0; i < 4; ++i) {
                int c[3];
                c[axis] = s; c[p] = cp[i]; c[q] = cq[i];
                val[i] = v[c[0]][c[1]][c[2]];
                in4[i] = val[i] > 0;
            }
            /* edge 0: along q at dp=0; edge 1: along p at dq=1; edge 2: along q at dp=1;
             * edge 3: along p at dq=0.  Edge ids need the (other two axes) coordinates in
             * increasing axis order. */
            for (int i = 0; i < 4; ++i) {
                int along = (i % 2 == 0) ? q : p;
                int c[3];
                c[axis] = s;
                if (along == q) c[p] = (i == 0) ? 0 : 1; else c[q] = (i == 1) ? 1 : 0;
                int o1 = along == 0 ? 1 : 0, o2 = along == 2 ? 1 : 2;
                fe[i] = mc_edge_id(along, c[o1], c[o2]);
            }
            int ncut = 0;
            for (int i = 0; i < 4; ++i) ncut += cut[fe[i]];
            if (ncut == 2) {
                int a = -1, b = -1;
                for (int i = 0; i < 4; ++i) if (cut[fe[i]]) { if (a < 0) a = fe[i]; else b = fe[i]; }
                MC_LINK(a, b);
            } else if (ncut == 4) {
                /* corner i touches face edges i-1 and i */
                double pin = 1.0, pout = 1.0;
                for (int i = 0; i < 4; ++i) { if (in4[i]) pin *= val[i]; else pout *= val[i]; }
                int cut_off_inside = !(pin > pout);
                for (int i = 0; i < 4; ++i)
                    if (in4[i] == cut_off_inside) MC_LINK(fe[(i + 3) % 4], fe[i]);
            }
        }
#undef MC_LINK
    int seen[12];
    for (int e = 0; e < 12; ++e) seen[e] = 0;
    double area = 0.0;
    for (int e0 = 0; e0 < 12; ++e0) {
        if (!cut[e0] || seen[e0] || nnb[e0] != 2) continue;
        seen[e0] = 1;
        int prev = e0, cur = nb[e0][0], last = -1;
        while (cur != e0) {
            seen[cur] = 1;
            if (last >= 0) {
                double a[3], b[3];
                for (int k = 0; k < 3; ++k) { a[k] = pos[last][k] - pos[e0][k]; b[k] = pos[cur][k] - pos[e0][k]; }
                double cx = a[1] * b[2] - a[2] * b[1];
                double cy = a[2] * b[0] - a[0] * b[2];
                double cz = a[0] * b[1] - a[1] * b[0];
                area += 0.5 * sqrt((cx * cx + cy * cy) + cz * cz);
            }
            last = cur;
            int nxt = nb[cur][0] == prev ? nb[cur][1] : nb[cur][0];
            prev = cur; cur = nxt;
        }
    }
    return area;
}

double orc_marching_cubes_area(int n0, int n1, int n2, const double *u, const double *dx) {
    double total = 0.0;
    for (long i = 0; i + 1 < n0; ++i)
        for (long j = 0; j + 1 < n1; ++j)
            for (long k = 0; k + 1 < n2; ++k) {
                double v[2][2][2];
                for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c)
                    v[a][b][c] = u[((i + a) * n1 + (j + b)) * n2 + (k + c)];
                total += mc_cube_area(v, dx);
            }
    return total;
}
