/*
 * TEST INFRASTRUCTURE — plain-C restatement of the merging-gym env hot path (float64).
 *
 * This is the checker and the timed CPU baseline, never the product: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may load it.  It follows, in the
 * reference's evaluation order,
 *   merging_gym/envs/merging_env.py:22-58    constants, lon2coord
 *   merging_gym/envs/merging_env.py:118-132  observe
 *   merging_gym/envs/merging_env.py:138-195  step
 *   merging_gym/envs/merging_env.py:198-206, 232-239  is_collided / corners
 *   merging_gym/envs/merging_env.py:208-230  reset
 *   scripts/helper.py:152-191                mpc_1d, as the closed form (vt - v0)/t of its QP
 * and mirrors oracle/merge_oracle.py::RefVecEnv (same auto-reset convention, same info
 * bit-field).  tests/test_oracle_c.py checks it against the NumPy oracle and the golden
 * traces recorded from the unmodified reference file.  Third-party boundaries (pygame Rect
 * truncation, Shapely intersects on touching rectangles, quadprog) are frozen by
 * definition — parity there is unpinned (see merge_oracle.py header).
 *
 * Build: see oracle/Makefile  (gcc -O2 -fno-fast-math -ffp-contract=off -fopenmp).
 * -ffp-contract=off matters: the reference rounds `v + acc*dT` and `p + v*dT` twice.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MGO_R 30000.0
#define MGO_H 1000.0
#define MGO_W 300.0
#define MGO_DT 0.2
#define MGO_RFIRST 2.0
#define MGO_RSECOND 1.0
#define MGO_RCOLLISION (-10.0)
#define MGO_VEL_PENALTY 0.001
#define MGO_START 50.0
#define MGO_END 950.0
#define MGO_PRED_T 3.0
#define MGO_TIME_LIMIT 500.0

#define INFO_COLLISION 0x01
#define INFO_WINNER_SHIFT 1
#define INFO_TIMEOUT 0x08
#define INFO_DONE 0x10
#define INFO_BAD_ACTION 0x80

/* stats vector layout (int64[10]); returns are carried separately as double[2] */
enum { ST_EPISODES, ST_COLLISIONS, ST_WINS_P1, ST_WINS_P2, ST_TIMEOUTS, ST_MERGES_OK,
       ST_SUM_LENGTH, ST_BAD_ACTIONS, ST_N };

typedef struct {
    double *pos1, *vel1, *pos2, *vel2, *ret1, *ret2, *time_stamp;
    int32_t *steps;
    uint8_t *winner, *done;
} MgoState;

static inline void lon2coord(double lon, int ego, double *x, double *y) {   /* :48-58 */
    double angle = atan2(MGO_H, MGO_R) - lon / MGO_R;
    *x = MGO_R * sin(angle);
    if (ego) *y = MGO_W / 2 + (MGO_R - MGO_R * cos(angle));
    else     *y = MGO_W / 2 - (MGO_R - MGO_R * cos(angle));
}

static inline void observe(double p1, double v1, double p2, double v2, double *o) {   /* :118-132 */
    double x1, y1, x2, y2;
    lon2coord(p1, 1, &x1, &y1);
    lon2coord(p2, 0, &x2, &y2);
    o[0] = x2 - x1; o[1] = y2 - y1; o[2] = v2 - v1; o[3] = MGO_END - p1; o[4] = v1;
    o[5] = x1 - x2; o[6] = y1 - y2; o[7] = v1 - v2; o[8] = MGO_END - p2; o[9] = v2;
}

static inline void reset_row(const MgoState *s, int64_t i) {                /* :208-230 */
    s->pos1[i] = MGO_START; s->vel1[i] = 20.0;
    s->pos2[i] = MGO_START; s->vel2[i] = 20.0;
    s->ret1[i] = 0.0; s->ret2[i] = 0.0; s->time_stamp[i] = 0.0;
    s->steps[i] = 0; s->winner[i] = 0; s->done[i] = 0;
}

void mgo_reset(const MgoState *s, int64_t n, const uint8_t *mask, double *obs) {
    for (int64_t i = 0; i < n; ++i) {
        if (!mask || mask[i]) reset_row(s, i);
        if (obs) observe(s->pos1[i], s->vel1[i], s->pos2[i], s->vel2[i], obs + 10 * i);
    }
}

/* One step of env i.  Returns the info byte. */
static inline uint8_t step_row(const MgoState *s, int64_t i, int a1, int a2, int pvp, int auto_reset,
                               double *obs, double *rew, double *term_obs, double *ep_ret,
                               int32_t *ep_len, int64_t *st, double *sumret) {
    int bad = (a1 < 0 || a1 > 4) || (pvp && (a2 < 0 || a2 > 4));
    if (a1 < 0) a1 = 0; if (a1 > 4) a1 = 4;
    if (a2 < 0) a2 = 0; if (a2 > 4) a2 = 4;

    s->time_stamp[i] += MGO_DT;                                              /* :141 */
    s->steps[i] += 1;
    int timeout = s->time_stamp[i] > MGO_TIME_LIMIT;                         /* :142 */
    int done = s->done[i] | timeout;

    double v1 = s->vel1[i], p1 = s->pos1[i], v2 = s->vel2[i], p2 = s->pos2[i];
    double acc1 = (10.0 * a1 - v1) / MGO_PRED_T;                             /* helper.py QP closed form */
    v1 = v1 + acc1 * MGO_DT; if (!(v1 > 0.0)) v1 = 0.0;                      /* :149 max(0, .) */
    p1 = p1 + v1 * MGO_DT;                                                   /* :150 */
    double acc2 = pvp ? (10.0 * a2 - v2) / MGO_PRED_T : 0.0;                 /* :152 */
    v2 = v2 + acc2 * MGO_DT; if (!(v2 > 0.0)) v2 = 0.0;
    p2 = p2 + v2 * MGO_DT;

    double x1, y1, x2, y2;
    lon2coord(p1, 1, &x1, &y1);
    lon2coord(p2, 0, &x2, &y2);
    double o[10] = { x2 - x1, y2 - y1, v2 - v1, MGO_END - p1, v1,
                     x1 - x2, y1 - y2, v1 - v2, MGO_END - p2, v2 };

    double r1 = 0.0 - MGO_VEL_PENALTY * fabs(v1 - 20.0);                     /* :158-159 */
    double r2 = 0.0 - MGO_VEL_PENALTY * fabs(v2 - 20.0);
    int w = s->winner[i];
    if (p1 > MGO_END) {                                                      /* :163-171 */
        if (w == 0) { w = 1; r1 += MGO_RFIRST; }
        else if (w == 1) r1 = 0.0;
        else { r1 += MGO_RSECOND; done = 1; }
    }
    if (p2 >= MGO_END) {                                                     /* :173-181 */
        if (w == 0) { w = 2; r2 += MGO_RFIRST; }
        else if (w == 2) r2 = 0.0;
        else { r2 += MGO_RSECOND; done = 1; }
    }
    /* :183-187, :198-206, :232-239 — integer Rects from C-truncated centres, closed-set overlap */
    long ty1 = (long)y1, tx1 = (long)x1, ty2 = (long)y2, tx2 = (long)x2;
    int col = (labs(ty1 - ty2) <= 4) && (labs(tx1 - tx2) <= 8);
    if (col) { done = 1; r1 += MGO_RCOLLISION; r2 += MGO_RCOLLISION; }

    double R1 = s->ret1[i] + r1, R2 = s->ret2[i] + r2;                       /* :191-192 */
    rew[0] = r1; rew[1] = r2;
    uint8_t info = (uint8_t)((col ? INFO_COLLISION : 0) | (w << INFO_WINNER_SHIFT) |
                             (timeout ? INFO_TIMEOUT : 0) | (done ? INFO_DONE : 0) |
                             (bad ? INFO_BAD_ACTION : 0));
    if (bad) st[ST_BAD_ACTIONS] += 1;

    if (done && !s->done[i]) {   /* done became true in this step (always so under auto-reset) */
        if (term_obs) memcpy(term_obs, o, sizeof o);
        if (ep_ret) { ep_ret[0] = R1; ep_ret[1] = R2; }
        if (ep_len) *ep_len = s->steps[i];
        st[ST_EPISODES] += 1; st[ST_COLLISIONS] += col; st[ST_WINS_P1] += (w == 1);
        st[ST_WINS_P2] += (w == 2); st[ST_TIMEOUTS] += timeout;
        st[ST_MERGES_OK] += (!col && !timeout); st[ST_SUM_LENGTH] += s->steps[i];
        sumret[0] += R1; sumret[1] += R2;
    }
    if (done && auto_reset) {
        reset_row(s, i);
        observe(s->pos1[i], s->vel1[i], s->pos2[i], s->vel2[i], obs);
    } else {
        s->pos1[i] = p1; s->vel1[i] = v1; s->pos2[i] = p2; s->vel2[i] = v2;
        s->ret1[i] = R1; s->ret2[i] = R2; s->winner[i] = (uint8_t)w; s->done[i] = (uint8_t)done;
        memcpy(obs, o, sizeof o);
    }
    return info;
}

/* One vector step.  a2 == NULL selects pve (`action2 is None`, :152).  Optional outputs may be
 * NULL.  stats: int64[ST_N] and sumret: double[2] are accumulated into (caller zeroes).
 * nthreads > 1 uses OpenMP with a static partition (per-thread partial stats, fixed order). */
void mgo_step(const MgoState *s, int64_t n, const uint8_t *a1, const uint8_t *a2, int auto_reset,
              double *obs, double *rew, uint8_t *done_out, uint8_t *info_out, double *term_obs,
              double *ep_ret, int32_t *ep_len, int64_t *stats, double *sumret, int nthreads) {
    int pvp = a2 != 0;
    if (nthreads < 1) nthreads = 1;
    int64_t tstats[64][ST_N];
    double tsum[64][2];
    if (nthreads > 64) nthreads = 64;
    memset(tstats, 0, sizeof tstats);
    memset(tsum, 0, sizeof tsum);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int t = 0; t < nthreads; ++t) {
        int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
        for (int64_t i = lo; i < hi; ++i) {
            uint8_t info = step_row(s, i, a1[i], pvp ? a2[i] : 0, pvp, auto_reset, obs + 10 * i,
                                    rew + 2 * i, term_obs ? term_obs + 10 * i : 0,
                                    ep_ret ? ep_ret + 2 * i : 0, ep_len ? ep_len + i : 0,
                                    tstats[t], tsum[t]);
            if (info_out) info_out[i] = info;
            if (done_out) done_out[i] = (info & INFO_DONE) ? 1 : 0;
        }
    }
    for (int t = 0; t < nthreads; ++t) {
        for (int k = 0; k < ST_N; ++k) stats[k] += tstats[t][k];
        sumret[0] += tsum[t][0]; sumret[1] += tsum[t][1];
    }
}

/* ---------------------------------------------------------------- Philox4x32-10 (Random123) */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void mgo_philox_actions(int64_t n, uint64_t seed, uint64_t env_id_base, uint64_t step,
                        uint8_t *a1, uint8_t *a2) {
    for (int64_t i = 0; i < n; ++i) {
        uint64_t id = env_id_base + (uint64_t)i;
        uint32_t c[4] = { (uint32_t)id, (uint32_t)(id >> 32), (uint32_t)step, (uint32_t)(step >> 32) };
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        a1[i] = (uint8_t)(((uint64_t)c[0] * 5u) >> 32);
        if (a2) a2[i] = (uint8_t)(((uint64_t)c[1] * 5u) >> 32);
    }
}

/* ---------------------------------------------------------------- arithmetic identities
 * The CUDA kernel replaces the two IEEE divisions on the path, x/3.0 and x/30000.0, by the
 * FMA sequence  q = x*y; r = fma(-d, q, x); q' = fma(r, y, q)  with y = RN(1/d)
 * (Markstein's correction step).  These helpers let the CPU tests confirm q' == x/d bit for
 * bit on the value ranges the env reaches.  Returns the number of mismatches. */
static inline double div_by_const(double x, double d, double y) {
    double q = x * y;
    double r = fma(-d, q, x);
    return fma(r, y, q);
}

int64_t mgo_check_div(const double *x, int64_t n, double d) {
    double y = 1.0 / d;
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        double a = x[i] / d, b = div_by_const(x[i], d, y);
        if (memcmp(&a, &b, sizeof a) != 0) ++bad;
    }
    return bad;
}

double mgo_atan2_h_r(void) { return atan2(MGO_H, MGO_R); }
int mgo_stats_len(void) { return ST_N; }
