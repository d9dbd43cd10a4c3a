/*
 * fmm_sched_model.c -- CPU model of the SCHEDULER of the device band rebuild (lsml_b200/csrc/fmm.cu).
 * TEST INFRASTRUCTURE ONLY (see oracle/lsml_oracle.c for the rules).
 *
 * oracle/fmm_replay_model.c checks that the replay operator V has the sequential marcher's state
 * as its fixed point.  fmm.cu does not sweep all candidates, though: it evaluates a candidate
 * only when it is QUEUED, and queues it (a) when it is created, (b) when a stencil voxel changed
 * and that voxel is, or was, within the candidate's `reach`, (c) not before the sweep its depth
 * hint from the previous rebuild names.  A flaw in (b) or (c) would leave a stale value behind
 * without any kernel noticing.  This file restates exactly those rules -- epochs, reach rounded
 * up to 16 bits, in-flux rule, depth levels in the state byte, hint tags and parking buckets,
 * sparse interface search from the previous band -- around the same V, sequentially, so that
 * tests/test_fmm_algorithm.py can run thousands of multi-rebuild sequences against the
 * sequential marcher on the CPU.  Two extreme visibility orders are modelled: in place (an
 * entry sees every earlier write of its sweep) and synchronous (it sees none); the GPU is
 * somewhere in between.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned char u8;
typedef uint16_t u16;
typedef uint32_t u32;

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
    double dx[3], idx2[3], band;
    const double *phi;
    double *T;
    u8 *state;      /* 0 far, 1 initial frozen, (level << 2) | 2 candidate */
    u32 *epoch;
    u16 *reach16;   /* top 16 bits of the float32 reach, rounded up */
    u16 *lvl;       /* (rebuild tag << 8) | depth: persists across rebuilds */
} sm_t;

static u16 reach_up16(float r) { u32 b; memcpy(&b, &r, 4); return (u16)((b + 0xffffu) >> 16); }
static float reach_from16(u16 h) { u32 b = (u32)h << 16; float f; memcpy(&f, &b, 4); return f; }

static int key_less(double ta, long ia, double tb, long ib) {
    float a = qmag(ta), b = qmag(tb);
    return (a < b) || (a == b && ia < ib);
}

/* same arithmetic as lsml_oracle.c:fmm_solve */
static double sm_solve(int ndim, const double *idx2, const double *v1in, const double *v2in,
                       double phi) {
    const double aa = 9.0 / 4.0, one_third = 1.0 / 3.0;
    double v1[3], v2[3];
    int active[3], nactive = 0;
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
                a += idx2[ax] * aa; b -= idx2[ax] * 2 * aa * tp; c += idx2[ax] * aa * (tp * tp);
            } else {
                a += idx2[ax]; b -= idx2[ax] * 2 * v1[ax]; c += idx2[ax] * (v1[ax] * v1[ax]);
            }
        }
        c -= 1;
        double det = b * b - 4 * a * c;
        if (det > 0) {
            double r = (phi > DBL_EPSILON) ? (-b + sqrt(det)) / 2.0 / a : (-b - sqrt(det)) / 2.0 / a;
            int causal = 1;
            for (int ax = 0; ax < ndim; ++ax) {
                if (!active[ax]) continue;
                if (qmag(r) <= qmag(v1[ax])) causal = 0;
                if (v2[ax] < DBL_MAX && !across(v1[ax], v2[ax]) && qmag(r) <= qmag(v2[ax])) causal = 0;
            }
            if (causal) return r;
        }
        int worst = -1, worst_is_v2 = 0; float wv = -1.0f;
        for (int ax = 0; ax < ndim; ++ax) {
            if (!active[ax]) continue;
            if (v2[ax] < DBL_MAX && !across(v1[ax], v2[ax]) && qmag(v2[ax]) > wv) { wv = qmag(v2[ax]); worst = ax; worst_is_v2 = 1; }
            if (qmag(v1[ax]) > wv) { wv = qmag(v1[ax]); worst = ax; worst_is_v2 = 0; }
        }
        if (worst_is_v2) v2[worst] = DBL_MAX; else { active[worst] = 0; nactive--; }
    }
    return 0.0;
}

static double stencil_update(const sm_t *m, double t1[3][2], double t2[3][2], int f1[3][2],
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
    return sm_solve(m->ndim, m->idx2, v1, v2, phi);
}

/* V(x) with the scheduler's by-products: reach (max of the last consumed event's magnitude and
 * the result's; +inf without a value) and level (1 + deepest consumed event; frozen voxels 0),
 * read from T / state through Tr / Sr (the synchronous mode passes snapshots). */
static double replay(const sm_t *m, const double *Tr, const u8 *Sr, long l, float *reach,
                     unsigned *level) {
    double t1[3][2], t2[3][2];
    int f1[3][2], f2[3][2], e1[3][2], e2[3][2], l1[3][2], l2[3][2];
    int any = 0;
    for (int ax = 0; ax < m->ndim; ++ax) {
        long c = (l / m->st[ax]) % m->shape[ax];
        for (int d = 0; d < 2; ++d) {
            int j = d ? 1 : -1;
            t1[ax][d] = t2[ax][d] = INFINITY;
            f1[ax][d] = f2[ax][d] = e1[ax][d] = e2[ax][d] = l1[ax][d] = l2[ax][d] = 0;
            if (c + j >= 0 && c + j < m->shape[ax]) {
                long n = l + j * m->st[ax];
                t1[ax][d] = Tr[n];
                f1[ax][d] = Sr[n] == 1;
                e1[ax][d] = (Sr[n] & 2) && Sr[n] != 1 && fabs(Tr[n]) <= m->band;
                l1[ax][d] = Sr[n] >> 2;
                any |= f1[ax][d];
            }
            if (c + 2 * j >= 0 && c + 2 * j < m->shape[ax]) {
                long n = l + 2 * j * m->st[ax];
                t2[ax][d] = Tr[n];
                f2[ax][d] = Sr[n] == 1;
                e2[ax][d] = (Sr[n] & 2) && Sr[n] != 1 && fabs(Tr[n]) <= m->band;
                l2[ax][d] = Sr[n] >> 2;
            }
        }
    }
    double phi = m->phi[l], t = INFINITY;
    int narrow = 0;
    float last_q = 0.f;
    unsigned deepest = 0;
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
        last_q = qmag(bt);
        unsigned lev = bw == 1 ? (unsigned)l1[bax][bd] : (unsigned)l2[bax][bd];
        if (lev > deepest) deepest = lev;
        int reeval;
        if (bw == 1) { e1[bax][bd] = 0; f1[bax][bd] = 1; reeval = 1; }
        else { e2[bax][bd] = 0; f2[bax][bd] = 1; reeval = f1[bax][bd] && narrow; }
        if (reeval) {
            double dn = stencil_update(m, t1, t2, f1, f2, phi);
            if (dn != 0.0) { t = dn; narrow = 1; }
        }
    }
    *reach = narrow ? fmaxf(last_q, qmag(t)) : INFINITY;
    *level = deepest + 1 > 63 ? 63 : deepest + 1;
    return narrow ? t : INFINITY;
}

static int init_distance(const sm_t *m, long i, double *d) {
    const double *phi = m->phi;
    if (phi[i] == 0.0) { *d = 0.0; return 1; }
    int borders = 0; double dsum = 0;
    for (int ax = 0; ax < m->ndim; ++ax) {
        long c = (i / m->st[ax]) % m->shape[ax];
        double ld = 0;
        for (int j = -1; j < 2; j += 2) {
            if (c + j < 0 || c + j >= m->shape[ax]) continue;
            double nv = phi[i + j * m->st[ax]];
            if (phi[i] * nv < 0) {
                borders = 1;
                double dd = m->dx[ax] * phi[i] / (phi[i] - nv);
                if (ld == 0 || ld > dd) ld = dd;
            }
        }
        if (ld > 0) dsum += 1 / ld / ld;
    }
    if (!borders) return 0;
    *d = phi[i] < 0 ? -sqrt(1 / dsum) : sqrt(1 / dsum);
    return 1;
}

typedef struct { long *v; long n, cap; } vec_t;
static void vpush(vec_t *q, long x) {
    if (q->n == q->cap) { q->cap = q->cap ? 2 * q->cap : 256; q->v = (long *)realloc(q->v, sizeof(long) * q->cap); }
    q->v[q->n++] = x;
}

typedef struct {
    sm_t *m;
    vec_t *ring_next, *buckets, *cands;
    u32 next;          /* epoch of the sweep being filled */
    u32 hint_tag;      /* (previous rebuild's tag) << 8 or ~0 */
    long last_bucket;
    long *hist_prev;   /* capacity check of the parking buckets, as sized on the device */
    long overflowed;
    int no_reach;      /* diagnostic: queue every stencil candidate of a changed voxel */
} sched_t;

static u32 target_sweep(const sched_t *s, u32 lvl) {
    u32 h = lvl & 0xffu;
    return ((lvl & 0xff00u) == s->hint_tag && h > s->next) ? h : s->next;
}

static void enqueue(sched_t *s, long n) {
    u32 target = target_sweep(s, s->m->lvl[n]);
    if (s->m->epoch[n] >= target) return;
    s->m->epoch[n] = target;
    if (target == s->next) vpush(s->ring_next, n);
    else {
        vpush(&s->buckets[target], n);
        if (s->buckets[target].n > s->hist_prev[target]) s->overflowed++;
        if ((long)target > s->last_bucket) s->last_bucket = target;
    }
}

static void expose_neighbours(sched_t *s, long l) {
    sm_t *m = s->m;
    for (int ax = 0; ax < m->ndim; ++ax) {
        long c = (l / m->st[ax]) % m->shape[ax];
        for (int j = -1; j < 2; j += 2) {
            if (c + j < 0 || c + j >= m->shape[ax]) continue;
            long n = l + j * m->st[ax];
            if (m->state[n] == 0) { m->state[n] = 2; vpush(s->cands, n); enqueue(s, n); }
        }
    }
}

static void publish_change(sched_t *s, long l, float q_min, int expose) {
    sm_t *m = s->m;
    u32 cur = s->next - 1;
    for (int ax = 0; ax < m->ndim; ++ax) {
        long c = (l / m->st[ax]) % m->shape[ax];
        for (int j = -2; j <= 2; ++j) {
            if (j == 0 || c + j < 0 || c + j >= m->shape[ax]) continue;
            long n = l + j * m->st[ax];
            u8 st = m->state[n];
            if (st != 1 && (st & 2)) {
                u32 target = target_sweep(s, m->lvl[n]);
                if (m->epoch[n] < target &&
                    (s->no_reach || m->epoch[n] == cur || q_min <= reach_from16(m->reach16[n])))
                    enqueue(s, n);
            } else if (expose && (j == 1 || j == -1) && st == 0) {
                m->state[n] = 2; vpush(s->cands, n); enqueue(s, n);
            }
        }
    }
}

/*
 * One rebuild.  Persistent arrays (caller-owned, size = number of voxels unless noted):
 *   lvl (u16), hist (long[2][256]), gen (long[1]).
 * prev_band: indices of the previous band (sparse interface search + hints allowed) or NULL
 * (dense search, no hints).  mode bit 0: synchronous sweeps; bit 1 (diagnostic): no reach
 * filter; bit 2 (diagnostic): no hints.  counters: long[4] = {sweeps,
 * replays, parked entries, bucket overflows}.  Returns sweeps, or -1 without convergence.
 */
/* fmm.cu:hint_tag_of -- tags 1..255, never 0: lvl == 0 (never evaluated / wiped) has no tag */
static u32 hint_tag_of(long gen) { return (u32)((gen % 255) + 1) << 8; }

long sm_rebuild(int ndim, const int *shape, const double *phi, const double *dx, double narrow,
                u16 *lvl, long *hist, long *gen, const long *prev_band, long n_prev, int mode,
                int max_sweeps, double *dist_out, u8 *frozen_out, long *counters) {
    sm_t m;
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
    m.epoch = (u32 *)calloc(m.size, sizeof(u32));
    m.reach16 = (u16 *)malloc(sizeof(u16) * m.size);
    m.lvl = lvl;
    for (long i = 0; i < m.size; ++i) { m.T[i] = INFINITY; m.reach16[i] = 0x7f80; }
    /* prepare: generation, bucket capacities from last rebuild's histogram */
    if ((gen[0] & 127) == 127) for (long i = 0; i < m.size; ++i) lvl[i] = 0;
    gen[0] += 1;
    const int cur = (int)(gen[0] & 1), prev = cur ^ 1;
    long *hist_cur = hist + 256 * cur, *hist_prev = hist + 256 * prev;
    for (int i = 0; i < 256; ++i) hist_cur[i] = 0;
    const int hints = prev_band != NULL && (gen[0] % 16) != 0 && !(mode & 4);
    const u32 gen_tag = hint_tag_of(gen[0]);
    vec_t ring_a = {0}, ring_b = {0}, cands = {0}, inits = {0};
    vec_t buckets[256];
    memset(buckets, 0, sizeof buckets);
    sched_t sc;
    sc.m = &m; sc.buckets = buckets; sc.cands = &cands; sc.last_bucket = 0;
    sc.hint_tag = hints ? hint_tag_of(gen[0] - 1) : 0xffffffffu;
    sc.hist_prev = hist_prev; sc.overflowed = 0; sc.no_reach = (mode & 2) != 0;
    /* initial frozen set */
    if (prev_band == NULL) {
        for (long i = 0; i < m.size; ++i) {
            double d;
            if (init_distance(&m, i, &d)) { m.T[i] = d; m.state[i] = 1; vpush(&inits, i); }
        }
    } else {
        for (long b = 0; b < n_prev; ++b) {
            long l = prev_band[b];
            double d;
            if (m.state[l] == 0 && init_distance(&m, l, &d)) { m.T[l] = d; m.state[l] = 1; vpush(&inits, l); }
            for (int ax = 0; ax < ndim; ++ax) {
                long c = (l / m.st[ax]) % m.shape[ax];
                for (int j = -1; j < 2; j += 2) {
                    if (c + j < 0 || c + j >= m.shape[ax]) continue;
                    long n = l + j * m.st[ax];
                    if (!(phi[l] * phi[n] < 0) || m.state[n] != 0) continue;
                    if (init_distance(&m, n, &d)) { m.T[n] = d; m.state[n] = 1; vpush(&inits, n); }
                }
            }
        }
    }
    /* seeds, queued for sweep 1 */
    vec_t *ring_cur = &ring_a, *ring_next = &ring_b;
    sc.next = 1; sc.ring_next = ring_cur;
    for (long t = 0; t < inits.n; ++t) expose_neighbours(&sc, inits.v[t]);
    double *Tsnap = NULL; u8 *Ssnap = NULL;
    if (mode & 1) { Tsnap = (double *)malloc(sizeof(double) * m.size); Ssnap = (u8 *)malloc(m.size); }
    long sweeps = 0, replays = 0, parked = 0;
    u32 sweep = 1;
    int conv = 0;
    double *new_T = NULL; float *new_reach = NULL; unsigned *new_level = NULL; long res_cap = 0;
    while (sweeps < max_sweeps) {
        vec_t *bk = sweep < 256 ? &buckets[sweep] : NULL;
        long work = ring_cur->n + (hints && bk ? bk->n : 0);
        if (work == 0) {
            if (!hints || (long)sweep >= sc.last_bucket) { conv = 1; break; }
            ++sweep;
            continue;
        }
        if (work > res_cap) {
            res_cap = 2 * work;
            new_T = (double *)realloc(new_T, sizeof(double) * res_cap);
            new_reach = (float *)realloc(new_reach, sizeof(float) * res_cap);
            new_level = (unsigned *)realloc(new_level, sizeof(unsigned) * res_cap);
        }
        if (hints && bk) parked += bk->n;
        sc.next = sweep + 1; sc.ring_next = ring_next;
        if (mode & 1) { memcpy(Tsnap, m.T, sizeof(double) * m.size); memcpy(Ssnap, m.state, m.size); }
        for (int phase = 0; phase < ((mode & 1) ? 2 : 1); ++phase) {
            for (long k = 0; k < work; ++k) {
                long l = k < ring_cur->n ? ring_cur->v[k] : bk->v[k - ring_cur->n];
                double nw; float reach; unsigned level;
                if ((mode & 1) && phase == 0) {
                    new_T[k] = replay(&m, Tsnap, Ssnap, l, &new_reach[k], &new_level[k]);
                    ++replays;
                    continue;
                }
                if (mode & 1) { nw = new_T[k]; reach = new_reach[k]; level = new_level[k]; }
                else { nw = replay(&m, m.T, m.state, l, &reach, &level); ++replays; }
                double old = m.T[l];
                m.state[l] = (u8)((level << 2) | 2);
                u32 lv = lvl[l];
                if (lv != (gen_tag | level)) {
                    if ((lv & 0xff00u) == gen_tag) hist_cur[lv & 0xff]--;
                    hist_cur[level]++;
                    lvl[l] = (u16)(gen_tag | level);
                }
                m.reach16[l] = reach_up16(reach);
                if (memcmp(&old, &nw, sizeof old) == 0) continue;
                m.T[l] = nw;
                float q_min = fminf(qmag(old), qmag(nw));
                publish_change(&sc, l, q_min, fabs(nw) <= m.band && !(fabs(old) <= m.band));
            }
        }
        ring_cur->n = 0;
        { vec_t *t = ring_cur; ring_cur = ring_next; ring_next = t; }
        ++sweep; ++sweeps;
    }
    for (long i = 0; i < m.size; ++i) {
        int fr = m.state[i] == 1 || ((m.state[i] & 2) && m.state[i] != 1 && fabs(m.T[i]) <= m.band);
        frozen_out[i] = (u8)fr;
        dist_out[i] = fr ? m.T[i] : 0.0;
    }
    /* invariant the device's bucket offsets rest on: hist_cur[h] == #voxels hinted (this tag, h) */
    {
        long recount[256];
        memset(recount, 0, sizeof recount);
        for (long i = 0; i < m.size; ++i)
            if ((lvl[i] & 0xff00u) == gen_tag) recount[lvl[i] & 0xff]++;
        for (int h = 0; h < 256; ++h)
            if (recount[h] != hist_cur[h]) sc.overflowed |= 2;
    }
    if (counters) { counters[0] = sweeps; counters[1] = replays; counters[2] = parked; counters[3] = sc.overflowed; }
    free(m.T); free(m.state); free(m.epoch); free(m.reach16);
    free(ring_a.v); free(ring_b.v); free(cands.v); free(inits.v);
    for (int i = 0; i < 256; ++i) free(buckets[i].v);
    free(Tsnap); free(Ssnap); free(new_T); free(new_reach); free(new_level);
    return conv ? sweeps : -1;
}
