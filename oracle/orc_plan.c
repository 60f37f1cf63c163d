/* TEST INFRASTRUCTURE — CPU oracle for the MoPA-RL hot path.  Not part of the product.
 *
 * RRT-Connect oracle: sequential restatement of what KinematicPlanner::plan does
 * (motion_planners/KinematicPlanner.cpp:125-251) through OMPL, which is NOT vendored in the
 * reference (README.md:66 clones ompl master, unpinned) -> PARITY UNPINNED against OMPL.
 * Restated from OMPL's published algorithm:
 *   - state space: one weight-1 subspace per active joint (mujoco_ompl_interface.cpp:149-281);
 *     R^1 with jnt_range bounds, SO(2) for unlimited hinges; distance = sum |dq_j|
 *   - ompl::geometric::RRTConnect::solve / growTree with range = `range`
 *   - DiscreteMotionValidator::checkMotion with resolution 0.005 (KinematicPlanner.cpp:87)
 *   - sentinels: invalid goal -> -5 (KinematicPlanner.cpp:181-184), no exact solution -> -4 (:249-250)
 * Two deliberate departures, shared with the CUDA planner: termination is an iteration cap
 * (`max_iter` main-loop iterations) instead of the wall-clock `timelimit`
 * (KinematicPlanner.cpp:188), and samples come from a counter-based generator keyed by
 * (seed, problem key, iteration, dimension) instead of OMPL's per-sampler mt19937 streams.
 */
#include <stdlib.h>
#include <string.h>

#include "orc_math.h"

int orc_valid_one(void *h, const R *q);
int orc_scene_nq(void *h);

#define ORC_MAXD 16

typedef struct {
    void *scene;
    int nq, nd;
    int adr[ORC_MAXD];
    R lo[ORC_MAXD], hi[ORC_MAXD], seg[ORC_MAXD];
    int so2[ORC_MAXD];
    R range;
    uint64_t seed;
    int max_nodes;
} orc_planner;

void *orc_planner_create(void *scene, const int32_t *active_qadr, const double *lo, const double *hi, const int32_t *is_so2,
                         int n_active, double range, double resolution, uint64_t seed, int max_nodes) {
    orc_planner *p = (orc_planner *)calloc(1, sizeof(orc_planner));
    p->scene = scene; p->nq = orc_scene_nq(scene); p->nd = n_active;
    for (int j = 0; j < n_active; j++) {
        p->adr[j] = active_qadr[j];
        p->lo[j] = (R)lo[j]; p->hi[j] = (R)hi[j]; p->so2[j] = is_so2[j];
        /* longest valid segment = resolution * maximum extent of the subspace (pi for SO2) */
        R ext = is_so2[j] ? RC(3.14159265358979323846) : (p->hi[j] - p->lo[j]);
        p->seg[j] = (R)resolution * ext;
    }
    p->range = (R)range; p->seed = seed; p->max_nodes = max_nodes;
    return p;
}
void orc_planner_destroy(void *h) { free(h); }

/* counter-based uniform in [0,1): splitmix64 finaliser over (seed, key, iter, dim) */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
static inline R urand(uint64_t seed, uint64_t key, uint32_t iter, uint32_t dim) {
    uint64_t x = mix64(seed ^ (key * 0x9E3779B97F4A7C15ULL));
    x = mix64(x + (((uint64_t)iter << 8) | dim) * 0xD1342543DE82EF95ULL);
    return (R)(uint32_t)(x >> 40) * RC(5.9604644775390625e-08); /* 24 bits * 2^-24 */
}

#define PI_R RC(3.14159265358979323846)
static inline R dist1(const orc_planner *p, int j, R a, R b) {
    R d = FABS_(a - b);
    if (p->so2[j] && d > PI_R) d = RC(2.0) * PI_R - d;
    return d;
}
static R distance(const orc_planner *p, const R *a, const R *b) {
    R d = 0;
    for (int j = 0; j < p->nd; j++) d = d + dist1(p, j, a[j], b[j]);
    return d;
}
static void interpolate(const orc_planner *p, const R *a, const R *b, R t, R *out) {
    for (int j = 0; j < p->nd; j++) {
        if (!p->so2[j]) { out[j] = MAD(b[j] - a[j], t, a[j]); continue; }
        R diff = b[j] - a[j];
        if (FABS_(diff) <= PI_R) out[j] = MAD(diff, t, a[j]);
        else {
            if (diff > 0) diff = RC(2.0) * PI_R - diff; else diff = -RC(2.0) * PI_R - diff;
            R v = MAD(-diff, t, a[j]);
            if (v > PI_R) v -= RC(2.0) * PI_R; else if (v < -PI_R) v += RC(2.0) * PI_R;
            out[j] = v;
        }
    }
}
static int seg_count(const orc_planner *p, const R *a, const R *b) {
    int n = 0;
    for (int j = 0; j < p->nd; j++) {
        int c = (int)ceil((double)(dist1(p, j, a[j], b[j]) / p->seg[j]));
        if (c > n) n = c;
    }
    return n;
}
static int state_valid(const orc_planner *p, const R *base_q, const R *x) {
    R q[128];
    memcpy(q, base_q, sizeof(R) * p->nq);
    for (int j = 0; j < p->nd; j++) q[p->adr[j]] = x[j];
    return orc_valid_one(p->scene, q);
}
/* DiscreteMotionValidator::checkMotion(s1, s2), s1 assumed valid; s2 checked by the caller */
static int check_interior(const orc_planner *p, const R *base_q, const R *s1, const R *s2) {
    int nd = seg_count(p, s1, s2);
    R x[ORC_MAXD];
    for (int mid = 1; mid < nd; mid++) { /* the bisection order of OMPL only affects early exit */
        interpolate(p, s1, s2, (R)mid / (R)nd, x);
        if (!state_valid(p, base_q, x)) return 0;
    }
    return 1;
}

typedef struct { R *x; int *parent; int n; } tree_t;
enum { TRAPPED = 0, ADVANCED = 1, REACHED = 2 };

static int nearest(const orc_planner *p, const tree_t *t, const R *x) {
    int best = 0;
    R bd = distance(p, t->x, x);
    for (int i = 1; i < t->n; i++) {
        R d = distance(p, t->x + (size_t)i * ORC_MAXD, x);
        if (d < bd) { bd = d; best = i; }
    }
    return best;
}
/* returns grow state; *added = index of the new node; xstate receives the state that was added */
static int grow(const orc_planner *p, const R *base_q, tree_t *t, int is_start_tree, const R *target, R *xstate, int *added) {
    int ni = nearest(p, t, target);
    const R *ns = t->x + (size_t)ni * ORC_MAXD;
    R d = distance(p, ns, target);
    R dstate[ORC_MAXD];
    int reach = 1;
    memcpy(dstate, target, sizeof(R) * p->nd);
    if (d > p->range) {
        interpolate(p, ns, target, p->range / d, dstate);
        int same = 1;
        for (int j = 0; j < p->nd; j++) if (dstate[j] != ns[j]) same = 0;
        if (same) return TRAPPED;
        reach = 0;
    }
    if (!state_valid(p, base_q, dstate)) return TRAPPED;
    int ok = is_start_tree ? check_interior(p, base_q, ns, dstate) : check_interior(p, base_q, dstate, ns);
    if (!ok) return TRAPPED;
    if (t->n >= p->max_nodes) return TRAPPED; /* tree storage exhausted: treated as no progress */
    memcpy(t->x + (size_t)t->n * ORC_MAXD, dstate, sizeof(R) * p->nd);
    t->parent[t->n] = ni;
    *added = t->n;
    t->n++;
    memcpy(xstate, dstate, sizeof(R) * p->nd);
    return reach ? REACHED : ADVANCED;
}

/* Plans one problem.  path: [max_path][nq] doubles; node_ids: [max_path] (goal-tree nodes have
 * bit 30 set).  Returns status (0 ok, -4 no exact solution, -5 invalid goal); *path_len rows
 * written; *iters = main-loop iterations used. */
int orc_plan(void *h, const double *start, const double *goal, uint64_t key, int max_iter, double *path, int32_t *node_ids,
             int max_path, int32_t *path_len, int32_t *iters, int32_t *n_nodes) {
    orc_planner *p = (orc_planner *)h;
    R base_q[128], s[ORC_MAXD], g[ORC_MAXD];
    for (int k = 0; k < p->nq; k++) base_q[k] = (R)start[k]; /* passive dims frozen at the start values */
    for (int j = 0; j < p->nd; j++) { s[j] = (R)start[p->adr[j]]; g[j] = (R)goal[p->adr[j]]; }
    *path_len = 0;
    if (iters) *iters = 0;
    if (n_nodes) { n_nodes[0] = 0; n_nodes[1] = 0; }
    if (!state_valid(p, base_q, g)) return -5;
    int in_bounds_s = 1, in_bounds_g = 1;
    for (int j = 0; j < p->nd; j++) {
        if (p->so2[j]) continue;
        if (s[j] > p->hi[j] || s[j] < p->lo[j]) in_bounds_s = 0;
        if (g[j] > p->hi[j] || g[j] < p->lo[j]) in_bounds_g = 0;
    }
    if (!in_bounds_s || !in_bounds_g || !state_valid(p, base_q, s)) return -4;
    tree_t T[2];
    for (int k = 0; k < 2; k++) {
        T[k].x = (R *)malloc(sizeof(R) * ORC_MAXD * p->max_nodes);
        T[k].parent = (int *)malloc(sizeof(int) * p->max_nodes);
        T[k].n = 1;
        T[k].parent[0] = -1;
    }
    memcpy(T[0].x, s, sizeof(R) * p->nd);
    memcpy(T[1].x, g, sizeof(R) * p->nd);
    int status = -4, start_tree = 1, it;
    int sm = -1, gm = -1;
    for (it = 0; it < max_iter; it++) {
        int ti = start_tree ? 0 : 1;
        int is_start = start_tree;
        start_tree = !start_tree;
        int oi = start_tree ? 0 : 1;
        R rstate[ORC_MAXD], xstate[ORC_MAXD];
        for (int j = 0; j < p->nd; j++) rstate[j] = MAD(p->hi[j] - p->lo[j], urand(p->seed, key, (uint32_t)it, (uint32_t)j), p->lo[j]);
        int added = -1, oadded = -1;
        int gs = grow(p, base_q, &T[ti], is_start, rstate, xstate, &added);
        if (gs == TRAPPED) continue;
        memcpy(rstate, xstate, sizeof(R) * p->nd);
        int gsc = grow(p, base_q, &T[oi], start_tree, rstate, xstate, &oadded);
        while (gsc == ADVANCED) gsc = grow(p, base_q, &T[oi], start_tree, rstate, xstate, &oadded);
        if (gsc == REACHED) {
            sm = start_tree ? oadded : added; /* node of the start tree */
            gm = start_tree ? added : oadded; /* node of the goal tree */
            status = 0;
            it++;
            break;
        }
    }
    if (iters) *iters = it;
    if (n_nodes) { n_nodes[0] = T[0].n; n_nodes[1] = T[1].n; }
    if (status == 0) {
        /* drop the duplicated junction state (OMPL steps one motion back on the start side if it can) */
        if (T[0].parent[sm] >= 0) sm = T[0].parent[sm]; else gm = T[1].parent[gm];
        int cnt = 0;
        for (int i = sm; i >= 0; i = T[0].parent[i]) cnt++;
        int n1 = cnt;
        for (int i = gm; i >= 0; i = T[1].parent[i]) cnt++;
        if (cnt > max_path) status = -4;
        else {
            int r = n1 - 1;
            for (int i = sm; i >= 0; i = T[0].parent[i], r--) {
                for (int k = 0; k < p->nq; k++) path[(size_t)r * p->nq + k] = (double)base_q[k];
                for (int j = 0; j < p->nd; j++) path[(size_t)r * p->nq + p->adr[j]] = (double)T[0].x[(size_t)i * ORC_MAXD + j];
                node_ids[r] = i;
            }
            r = n1;
            for (int i = gm; i >= 0; i = T[1].parent[i], r++) {
                for (int k = 0; k < p->nq; k++) path[(size_t)r * p->nq + k] = (double)base_q[k];
                for (int j = 0; j < p->nd; j++) path[(size_t)r * p->nq + p->adr[j]] = (double)T[1].x[(size_t)i * ORC_MAXD + j];
                node_ids[r] = i | (1 << 30);
            }
            *path_len = cnt;
        }
    }
    for (int k = 0; k < 2; k++) { free(T[k].x); free(T[k].parent); }
    return status;
}
