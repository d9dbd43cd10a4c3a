/*
 * skfmm_literal.c -- "upstream-literal" restatement of scikit-fmm's distance marcher
 * (TEST INFRASTRUCTURE; only tests/, tools/ reports and bench.py's CPU legs may load it).
 *
 * The reference's band is the output of `skfmm.distance(arr, narrow=band, dx=dx)`
 * (/root/reference/lsml/util/distance_transform.py:49).  scikit-fmm is a third-party package
 * (PyPI scikit_fmm, unpinned in /root/reference/setup.py:59), it is not installed in this image
 * and its source is not under /root/reference, so it cannot be compiled or imported here.
 *
 * oracle/lsml_oracle.c:orc_fmm_distance -- the marcher the product is bit-exact to -- departs
 * from upstream ON PURPOSE at four points (float32-rounded ordering keys with a flat-index
 * tie-break, a causality test with contributor eviction, no `break` at the band edge, value2
 * reset).  This file is the other side of that comparison: scikit-fmm's C++ restated as
 * literally as memory of its source allows (src/heap.cpp, src/base_marcher.cpp,
 * src/distance_marcher.cpp of the 2019-2022 releases), WITHOUT any of those four changes:
 *
 *   heap          the array heap of heap.cpp: push = append + _siftDown(0, last); pop = move the
 *                 last entry to the root + _siftUp(0); set = _siftUp if the key grew, then
 *                 _siftDown(0, pos) (the CPython heapq algorithm, whose names it keeps: "siftUp"
 *                 walks DOWN to a leaf promoting the smaller child -- right child on ties -- and
 *                 then bubbles the item back up).  Keys are float64 |d|; ties are resolved by
 *                 these mechanics and nothing else.
 *   update        distanceMarcher::updatePointOrderTwo: per axis, both directions in the order
 *                 -1, +1; a frozen direct neighbour with fabs(d) < fabs(value1) becomes value1;
 *                 value2 is (re)assigned only when the voxel two cells away qualifies -- a value2
 *                 found for the first direction SURVIVES when the second direction supplies the
 *                 smaller value1 without a qualifying second neighbour (upstream quirk, kept).
 *   quadratic     solveQuadratic: c -= 1; det = b^2 - 4ac; det >= 0: root by the sign of phi
 *                 (phi > DBL_EPSILON takes +sqrt).  det < 0: the releases since 2019 throw
 *                 "Negative discriminant in distance marcher quadratic" and updatePointOrderTwo
 *                 catches it and returns "the position of the minimum as a more reasonable
 *                 approximation", -b / (2a); older releases returned 0 there (the caller then
 *                 leaves the voxel alone -- except initalizeNarrow, which pushes the 0 and so
 *                 freezes a spurious zero).  orc_skfmm_literal_set_old_det(1) selects the older
 *                 behaviour.  No causality test of any kind in either.
 *   march         baseMarcher::solve: pop; `if (narrow != 0 && fabs(value) > narrow) break;`
 *                 freeze; update the four-per-axis stencil exactly in upstream's order (direct
 *                 neighbour, then the "jump over a frozen point" second-order update, per
 *                 direction, per axis).
 *   output        every voxel that was never frozen is masked (pinned by the reference's own
 *                 known answer, lsml/util/test/test_distance_transform.py:11-21: the two end
 *                 voxels are tentative at -1.5 when the march stops and come back masked).
 *
 * What "from memory" means: nothing here could be diffed against the real source.  The
 * reference's known answers pass (tests/test_skfmm_literal.py); beyond them this file is a
 * MEASURING STICK for how far the product's marcher can be from upstream (tools/
 * band_gap_report.py -> profiles/r2_band_gap_vs_upstream_literal.json), not a pin.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char u8;

enum { L_FAR = 0, L_NARROW = 1, L_FROZEN = 2 };

typedef struct {
    /* heap.cpp: heap_ holds list indices; distance_/address_/backPointer_ are per list index */
    long max_len, heap_len, list_len;
    double *distance;
    long *back_pointer, *heap, *address;
} lheap_t;

static void lheap_sift_down(lheap_t *h, long start_pos, long pos) {
    long new_item = h->heap[pos];
    while (pos > start_pos) {
        long parent_pos = (pos - 1) >> 1;
        long parent = h->heap[parent_pos];
        if (h->distance[new_item] < h->distance[parent]) {
            h->heap[pos] = parent;
            h->back_pointer[parent] = pos;
            pos = parent_pos;
            continue;
        }
        break;
    }
    h->heap[pos] = new_item;
    h->back_pointer[new_item] = pos;
}

static void lheap_sift_up(lheap_t *h, long pos) {
    long end_pos = h->heap_len, start_pos = pos;
    long new_item = h->heap[pos];
    long child_pos = 2 * pos + 1;
    while (child_pos < end_pos) {
        long right_pos = child_pos + 1;
        if (right_pos < end_pos &&
            !(h->distance[h->heap[child_pos]] < h->distance[h->heap[right_pos]]))
            child_pos = right_pos;
        h->heap[pos] = h->heap[child_pos];
        h->back_pointer[h->heap[child_pos]] = pos;
        pos = child_pos;
        child_pos = 2 * pos + 1;
    }
    h->heap[pos] = new_item;
    lheap_sift_down(h, start_pos, pos);
}

static long lheap_push(lheap_t *h, long address, double value) {
    h->heap[h->heap_len] = h->list_len;
    h->address[h->list_len] = address;
    h->distance[h->list_len] = value;
    h->back_pointer[h->list_len] = h->heap_len;
    h->heap_len++;
    h->list_len++;
    lheap_sift_down(h, 0, h->heap_len - 1);
    return h->list_len - 1;
}

static void lheap_pop(lheap_t *h, long *address, double *value) {
    long loc = h->heap[0];
    *value = h->distance[loc];
    *address = h->address[loc];
    h->heap[0] = h->heap[h->heap_len - 1];
    h->back_pointer[h->heap[0]] = 0;
    h->heap_len--;
    if (h->heap_len > 0) lheap_sift_up(h, 0);
}

static void lheap_set(lheap_t *h, long index, double new_distance) {
    double old_distance = h->distance[index];
    long pos = h->back_pointer[index];
    h->distance[index] = new_distance;
    if (new_distance > old_distance) lheap_sift_up(h, pos);
    if (h->distance[h->heap[pos]] != new_distance) return;
    lheap_sift_down(h, 0, pos);
}

typedef struct {
    int ndim;
    int shape[3];
    long shift[3];
    long size;
    double idx2[3];
    const double *phi;
    double *dist;
    signed char *flag;
    long *heapptr;
    lheap_t heap;
} lm_t;

/* baseMarcher::_getN without periodic boundaries: the address `dir` cells along `dim`, or -1 when
 * it is outside the grid or carries `flag` */
static long get_n(const lm_t *m, long current, int dim, int dir, int flag) {
    long coord = (current / m->shift[dim]) % m->shape[dim];
    long newc = coord + dir;
    if (newc >= m->shape[dim] || newc < 0) return -1;
    long newa = current + (long)dir * m->shift[dim];
    if (m->flag[newa] == flag) return -1;
    return newa;
}

#define L_MASK 3 /* no voxel is ever masked on this path */

static int g_old_det = 0;
void orc_skfmm_literal_set_old_det(int on) { g_old_det = on; }

/* distanceMarcher::updatePointOrderTwo + solveQuadratic */
static double update_point(const lm_t *m, long i) {
    const double aa = 9.0 / 4.0, one_third = 1.0 / 3.0;
    double a = 0, b = 0, c = 0;
    for (int dim = 0; dim < m->ndim; ++dim) {
        double value1 = DBL_MAX, value2 = DBL_MAX;
        for (int j = -1; j < 2; j += 2) {
            long naddr = get_n(m, i, dim, j, L_MASK);
            if (naddr != -1 && m->flag[naddr] == L_FROZEN) {
                if (fabs(m->dist[naddr]) < fabs(value1)) {
                    value1 = m->dist[naddr];
                    long naddr2 = get_n(m, i, dim, j * 2, L_MASK);
                    if (naddr2 != -1 && m->flag[naddr2] == L_FROZEN &&
                        ((m->dist[naddr2] <= value1 && value1 >= 0) ||
                         (m->dist[naddr2] >= value1 && value1 <= 0)))
                        value2 = m->dist[naddr2];
                }
            }
        }
        if (value2 < DBL_MAX) {
            double tp = one_third * (4 * value1 - value2);
            a += m->idx2[dim] * aa;
            b -= m->idx2[dim] * 2 * aa * tp;
            c += m->idx2[dim] * aa * (tp * tp);
        } else if (value1 < DBL_MAX) {
            a += m->idx2[dim];
            b -= m->idx2[dim] * 2 * value1;
            c += m->idx2[dim] * (value1 * value1);
        }
    }
    c -= 1;
    double det = b * b - 4 * a * c;
    if (g_old_det ? det > 0 : det >= 0) {
        if (m->phi[i] > DBL_EPSILON) return (-b + sqrt(det)) / 2.0 / a;
        return (-b - sqrt(det)) / 2.0 / a;
    }
    if (g_old_det || a == 0) return 0.0;
    return -b / (2.0 * a);
}

/* Same contract as orc_fmm_distance: dist_out = signed distance (0 where never frozen),
 * frozen_out = ~mask; returns the number of frozen voxels or -1 without any zero contour. */
long orc_skfmm_literal_distance(int ndim, const int *shape, const double *phi, const double *dx,
                                double narrow, double *dist_out, u8 *frozen_out) {
    lm_t m;
    memset(&m, 0, sizeof m);
    m.ndim = ndim;
    m.size = 1;
    for (int a = 0; a < ndim; ++a) { m.shape[a] = shape[a]; m.size *= shape[a]; }
    { long s = 1; for (int a = ndim - 1; a >= 0; --a) { m.shift[a] = s; s *= shape[a]; } }
    for (int a = 0; a < ndim; ++a) m.idx2[a] = 1.0 / dx[a] / dx[a];
    m.phi = phi;
    m.dist = dist_out;
    m.flag = (signed char *)calloc(m.size, 1);
    m.heapptr = (long *)malloc(sizeof(long) * m.size);
    for (long i = 0; i < m.size; ++i) m.dist[i] = 0.0;

    /* initalizeFrozen */
    long nfrozen = 0;
    for (long i = 0; i < m.size; ++i)
        if (phi[i] == 0.0) { m.flag[i] = L_FROZEN; m.dist[i] = 0.0; nfrozen++; }
    for (long i = 0; i < m.size; ++i) {
        if (m.flag[i] != L_FAR) continue;
        double ldistance[3];
        int borders = 0;
        for (int dim = 0; dim < ndim; ++dim) {
            ldistance[dim] = 0;
            for (int j = -1; j < 2; j += 2) {
                long naddr = get_n(&m, i, dim, j, L_MASK);
                if (naddr != -1 && phi[i] * phi[naddr] < 0) {
                    borders = 1;
                    double d = dx[dim] * phi[i] / (phi[i] - phi[naddr]);
                    if (ldistance[dim] == 0 || ldistance[dim] > d) ldistance[dim] = d;
                }
            }
        }
        if (borders) {
            double dsum = 0;
            for (int dim = 0; dim < ndim; ++dim)
                if (ldistance[dim] > 0) dsum += 1 / ldistance[dim] / ldistance[dim];
            m.dist[i] = (phi[i] < 0) ? -sqrt(1 / dsum) : sqrt(1 / dsum);
            m.flag[i] = L_FROZEN;
            nfrozen++;
        }
    }
    if (nfrozen == 0) { free(m.flag); free(m.heapptr); return -1; }

    /* march(): heap sized by the number of far points */
    long max_heap = 0;
    for (long i = 0; i < m.size; ++i) if (m.flag[i] == L_FAR) max_heap++;
    if (max_heap == 0) max_heap = 1;
    m.heap.max_len = max_heap;
    m.heap.distance = (double *)malloc(sizeof(double) * max_heap);
    m.heap.back_pointer = (long *)malloc(sizeof(long) * max_heap);
    m.heap.heap = (long *)malloc(sizeof(long) * max_heap);
    m.heap.address = (long *)malloc(sizeof(long) * max_heap);

    /* initalizeNarrow */
    for (long i = 0; i < m.size; ++i) {
        if (m.flag[i] != L_FAR) continue;
        for (int dim = 0; dim < ndim; ++dim)
            for (int j = -1; j < 2; j += 2) {
                long naddr = get_n(&m, i, dim, j, L_MASK);
                if (naddr != -1 && m.flag[naddr] == L_FROZEN && m.flag[i] == L_FAR) {
                    m.flag[i] = L_NARROW;
                    double d = update_point(&m, i);
                    m.dist[i] = d;
                    m.heapptr[i] = lheap_push(&m.heap, i, fabs(d));
                }
            }
    }

    /* solve() */
    while (m.heap.heap_len > 0) {
        double value = 0;
        long addr = 0;
        lheap_pop(&m.heap, &addr, &value);
        if (narrow != 0 && fabs(value) > narrow) break;
        m.flag[addr] = L_FROZEN;
        nfrozen++;
        for (int dim = 0; dim < ndim; ++dim)
            for (int j = -1; j < 2; j += 2) {
                long naddr = get_n(&m, addr, dim, j, L_FROZEN);
                if (naddr != -1 && m.flag[naddr] != L_FROZEN) {
                    if (m.flag[naddr] == L_NARROW) {
                        double d = update_point(&m, naddr);
                        if (d) {
                            lheap_set(&m.heap, m.heapptr[naddr], fabs(d));
                            m.dist[naddr] = d;
                        }
                    } else if (m.flag[naddr] == L_FAR) {
                        double d = update_point(&m, naddr);
                        if (d) {
                            m.dist[naddr] = d;
                            m.flag[naddr] = L_NARROW;
                            m.heapptr[naddr] = lheap_push(&m.heap, naddr, fabs(d));
                        }
                    }
                }
                /* order 2: update the point two cells away, "jumping" over a frozen point */
                long local_naddr = get_n(&m, addr, dim, j, L_MASK);
                if (local_naddr != -1 && m.flag[local_naddr] == L_FROZEN) {
                    long naddr2 = get_n(&m, addr, dim, j * 2, L_FROZEN);
                    if (naddr2 != -1 && m.flag[naddr2] == L_NARROW) {
                        double d = update_point(&m, naddr2);
                        if (d) {
                            lheap_set(&m.heap, m.heapptr[naddr2], fabs(d));
                            m.dist[naddr2] = d;
                        }
                    }
                }
            }
    }
    for (long i = 0; i < m.size; ++i) {
        int fr = (m.flag[i] == L_FROZEN);
        frozen_out[i] = (u8)fr;
        if (!fr) m.dist[i] = 0.0;
    }
    free(m.flag); free(m.heapptr);
    free(m.heap.distance); free(m.heap.back_pointer); free(m.heap.heap); free(m.heap.address);
    return nfrozen;
}
