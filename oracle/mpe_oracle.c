/*
 * mpe_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker; never linked into, imported by or
 * executed from the product path).
 *
 * A plain-C, one-world-at-a-time restatement of the reference algorithm on the hot path
 * (openai/multiagent-particle-envs @ 83ba4d1).  Each function cites the reference file:line it
 * follows.  Compiled twice by oracle/Makefile:
 *     -DREAL=double -DSFX=_f64   the reference's own arithmetic (NumPy float64); pinned against
 *                                 the imported Python reference via the tests/golden fixtures
 *     -DREAL=float  -DSFX=_f32   same operation order in fp32 (built with -ffp-contract=off):
 *                                 reproduces the CUDA kernels' collision flags bit-for-bit when fed
 *                                 the kernels' stored fp32 state
 *
 * Layout (array-of-structs per world, the natural host layout):
 *   pv   REAL [n][A][4]      (p_pos.x, p_pos.y, p_vel.x, p_vel.y)
 *   lm   REAL [n][L][2]
 *   comm REAL [n][A][dim_c]  state.c of every agent
 *   goal int32 [n][G]
 *   act  REAL [n][sum_i act_dim_i], obs REAL [n][sum_i obs_dim_i], rew REAL [n][A],
 *   done uint8 [n][A], info REAL [n][A][info_dim]
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/mpe_b200.h"

#ifndef REAL
#define REAL double
#define SFX _f64
#endif
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

typedef REAL real;
#define R(x) ((real)(x))

static inline real r_sqrt(real x) { return sizeof(real) == 4 ? (real)sqrtf((float)x) : (real)sqrt((double)x); }
static inline real r_exp(real x) { return sizeof(real) == 4 ? (real)expf((float)x) : (real)exp((double)x); }
static inline real r_log1p(real x) { return sizeof(real) == 4 ? (real)log1pf((float)x) : (real)log1p((double)x); }
static inline real r_abs(real x) { return x < 0 ? -x : x; }

#define MAXA MPE_MAX_AGENTS
#define MAXL MPE_MAX_LANDMARKS
#define MAXC 16

/* ------------------------------------------------------------------------------------------ */
/* shapes                                                                                      */

int FN(mpe_oracle_act_dim)(const mpe_desc *d, int i) {
    /* environment.py:45-63: Discrete(2*dim_p+1) if movable, Discrete(dim_c) if not silent */
    return (d->agent_movable[i] ? 5 : 0) + (d->agent_silent[i] ? 0 : d->dim_c);
}

int FN(mpe_oracle_obs_dim)(const mpe_desc *d, int i) {
    const int A = d->n_agents, L = d->n_landmarks;
    switch (d->scenario) {
    case MPE_SCN_SIMPLE: return 2 + 2 * L;                               /* simple.py:45-50 */
    case MPE_SCN_SPREAD: return 4 + 2 * L + 2 * (A - 1) + d->dim_c * (A - 1); /* simple_spread.py:84-100 */
    case MPE_SCN_TAG: {                                                  /* simple_tag.py:131-147 */
        int n_good_others = (A - d->n_adversaries) - (d->agent_adversary[i] ? 0 : 1);
        return 4 + 2 * L + 2 * (A - 1) + 2 * n_good_others;
    }
    case MPE_SCN_WORLD_COMM: {                                           /* simple_world_comm.py:224-287 */
        int n_good_others = (A - d->n_adversaries) - (d->agent_adversary[i] ? 0 : 1);
        int base = 4 + 2 * L + 2 * (A - 1) + 2 * n_good_others + 2;
        return d->agent_adversary[i] ? base + d->dim_c : base;
    }
    case MPE_SCN_ADVERSARY: return d->agent_adversary[i] ? 2 * L + 2 * (A - 1) : 2 + 2 * L + 2 * (A - 1);  /* simple_adversary.py:121-139 */
    case MPE_SCN_PUSH: return d->agent_adversary[i] ? 2 + 2 * L + 2 * (A - 1) : 2 + 2 + 3 + 2 * L + 3 * L + 2 * (A - 1);  /* simple_push.py:76-96 */
    case MPE_SCN_SPEAKER_LISTENER: return d->agent_movable[i] ? 2 + 2 * L + d->dim_c : 3;  /* simple_speaker_listener.py:69-92 */
    case MPE_SCN_REFERENCE: return 2 + 2 * L + 3 + d->dim_c * (A - 1);   /* simple_reference.py:61-80 */
    case MPE_SCN_CRYPTO: return i == 2 ? 2 * d->dim_c : (d->agent_adversary[i] ? d->dim_c : 2 * d->dim_c);  /* simple_crypto.py:124-169 */
    default: return -1;
    }
}

int FN(mpe_oracle_num_goals)(const mpe_desc *d) {
    switch (d->scenario) {
    case MPE_SCN_ADVERSARY: case MPE_SCN_PUSH: case MPE_SCN_SPEAKER_LISTENER: return 1;
    case MPE_SCN_REFERENCE: case MPE_SCN_CRYPTO: return 2;
    default: return 0;
    }
}

int FN(mpe_oracle_info_dim)(const mpe_desc *d) {
    switch (d->scenario) {
    case MPE_SCN_SIMPLE: return 0;
    case MPE_SCN_SPREAD: return 4;      /* simple_spread.py:47-63 */
    case MPE_SCN_TAG: return 1;         /* simple_tag.py:57-66   */
    case MPE_SCN_WORLD_COMM: return 1;  /* simple_world_comm.py:115-123 */
    case MPE_SCN_ADVERSARY: return d->n_landmarks + 1;   /* simple_adversary.py:61-70 */
    case MPE_SCN_CRYPTO: return 2 * d->dim_c;            /* simple_crypto.py:66-67 */
    case MPE_SCN_PUSH: case MPE_SCN_SPEAKER_LISTENER: case MPE_SCN_REFERENCE: return 0;
    default: return -1;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* MultiAgentEnv._set_action (environment.py:144-192) for every agent of one world             */

static void set_action_world(const mpe_desc *d, const real *act, uint32_t flags, real *u, real *c) {
    const int A = d->n_agents, C = d->dim_c;
    const real *a = act;
    for (int i = 0; i < A; ++i) {
        real ux = 0, uy = 0;                                   /* :145 */
        for (int k = 0; k < C; ++k) c[i * C + k] = 0;          /* :146 */
        if (d->agent_movable[i]) {
            if (flags & MPE_FLAG_DISCRETE_ACTION_INPUT) {      /* :161-167 */
                int v = (int)a[0];
                if (v == 1) ux = R(-1.0);
                if (v == 2) ux = R(+1.0);
                if (v == 3) uy = R(-1.0);
                if (v == 4) uy = R(+1.0);
                a += 1;
            } else {
                real p[5] = {a[0], a[1], a[2], a[3], a[4]};
                if (flags & MPE_FLAG_FORCE_DISCRETE_ACTION) {  /* :169-172 (np.argmax: first max) */
                    int best = 0;
                    for (int k = 1; k < 5; ++k) if (p[k] > p[best]) best = k;
                    for (int k = 0; k < 5; ++k) p[k] = (k == best) ? R(1.0) : R(0.0);
                }
                ux += p[1] - p[2];                             /* :174 */
                uy += p[3] - p[4];                             /* :175 */
                a += 5;
            }
            real sens = R(d->agent_sens[i]);                   /* :178-181 */
            ux *= sens;
            uy *= sens;
        }
        if (!d->agent_silent[i]) {                             /* :183-190 */
            if (flags & MPE_FLAG_DISCRETE_ACTION_INPUT) {
                int v = (int)a[0];
                if (v >= 0 && v < C) c[i * C + v] = R(1.0);
                a += 1;
            } else {
                for (int k = 0; k < C; ++k) c[i * C + k] = a[k];
                a += C;
            }
        }
        u[i * 2 + 0] = ux;
        u[i * 2 + 1] = uy;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* World.step (core.py:117-131)                                                                */

static void world_step_world(const mpe_desc *d, real *pv, const real *lm, real *comm, const real *u,
                             const real *c) {
    const int A = d->n_agents, L = d->n_landmarks, E = A + L, C = d->dim_c;
    real fx[MAXA + MAXL], fy[MAXA + MAXL];
    int has_f[MAXA + MAXL];
    real px[MAXA + MAXL], py[MAXA + MAXL], size[MAXA + MAXL];
    int collide[MAXA + MAXL], movable[MAXA + MAXL];
    for (int e = 0; e < E; ++e) {
        has_f[e] = 0; fx[e] = 0; fy[e] = 0;
        if (e < A) {
            px[e] = pv[e * 4 + 0]; py[e] = pv[e * 4 + 1];
            size[e] = R(d->agent_size[e]); collide[e] = d->agent_collide[e]; movable[e] = d->agent_movable[e];
        } else {
            px[e] = lm[(e - A) * 2 + 0]; py[e] = lm[(e - A) * 2 + 1];
            size[e] = R(d->landmark_size[e - A]); collide[e] = d->landmark_collide[e - A]; movable[e] = 0;
        }
    }
    /* apply_action_force (core.py:134-140); u_noise is None everywhere */
    for (int i = 0; i < A; ++i)
        if (movable[i]) { fx[i] = u[i * 2 + 0]; fy[i] = u[i * 2 + 1]; has_f[i] = 1; }
    /* apply_environment_force (core.py:143-155): upper triangle in (a, b) lexicographic order */
    const real k = R(d->contact_margin), cf = R(d->contact_force);
    for (int a = 0; a < E; ++a)
        for (int b = a + 1; b < E; ++b) {
            /* get_collision_force (core.py:180-196) */
            if (!collide[a] || !collide[b]) continue;                      /* :181-182 */
            real dx = px[a] - px[b], dy = py[a] - py[b];                    /* :186 */
            real dist = r_sqrt(dx * dx + dy * dy);                          /* :187 */
            real dist_min = size[a] + size[b];                              /* :189 */
            real x = -(dist - dist_min) / k;                                /* :192 argument */
            /* np.logaddexp(0, x) = max(x, 0) + log1p(exp(-|x|)) */
            real pen = ((x > 0 ? x : 0) + r_log1p(r_exp(-r_abs(x)))) * k;   /* :192 */
            real f_x = cf * dx / dist * pen;                                /* :193 */
            real f_y = cf * dy / dist * pen;
            if (movable[a]) { fx[a] = +f_x + fx[a]; fy[a] = +f_y + fy[a]; has_f[a] = 1; }   /* :194,149-151 */
            if (movable[b]) { fx[b] = -f_x + fx[b]; fy[b] = -f_y + fy[b]; has_f[b] = 1; }   /* :195,152-154 */
        }
    /* integrate_state (core.py:158-169); only agents are movable in the reference scenarios */
    const real dt = R(d->dt), keep = R(1.0 - d->damping);
    for (int i = 0; i < A; ++i) {
        if (!movable[i]) continue;                                          /* :160 */
        real vx = pv[i * 4 + 2] * keep, vy = pv[i * 4 + 3] * keep;          /* :161 */
        if (has_f[i]) {                                                     /* :162-163 */
            real m = R(d->agent_mass[i]);
            vx += (fx[i] / m) * dt;
            vy += (fy[i] / m) * dt;
        }
        if (d->agent_max_speed[i] >= 0) {                                   /* :164-168 */
            real ms = R(d->agent_max_speed[i]);
            real speed = r_sqrt(vx * vx + vy * vy);
            if (speed > ms) { vx = vx / speed * ms; vy = vy / speed * ms; }
        }
        pv[i * 4 + 0] += vx * dt;                                           /* :169 */
        pv[i * 4 + 1] += vy * dt;
        pv[i * 4 + 2] = vx;
        pv[i * 4 + 3] = vy;
    }
    /* update_agent_state (core.py:171-177); c_noise is None everywhere */
    for (int i = 0; i < A; ++i)
        for (int q = 0; q < C; ++q) comm[i * C + q] = d->agent_silent[i] ? R(0.0) : c[i * C + q];
}

/* ------------------------------------------------------------------------------------------ */
/* scenario helpers                                                                            */

static inline real dist2d(real ax, real ay, real bx, real by) {
    real dx = ax - bx, dy = ay - by;
    return r_sqrt(dx * dx + dy * dy);   /* np.sqrt(np.sum(np.square(delta_pos))) */
}

/* is_collision (simple_spread.py:66-70, simple_tag.py:68-72, simple_world_comm.py:126-130) */
static inline int is_collision(real ax, real ay, real sa, real bx, real by, real sb) {
    return dist2d(ax, ay, bx, by) < sa + sb;
}

/* bound() (simple_tag.py:103-108, simple_world_comm.py:170-175) */
static inline real bound_pen(real x) {
    if (x < R(0.9)) return 0;
    if (x < R(1.0)) return (x - R(0.9)) * R(10.0);
    real e = r_exp(R(2.0) * x - R(2.0));
    return e < R(10.0) ? e : R(10.0);
}

#define PX(i) pv[(i) * 4 + 0]
#define PY(i) pv[(i) * 4 + 1]
#define VX(i) pv[(i) * 4 + 2]
#define VY(i) pv[(i) * 4 + 3]
#define LX(l) lm[(l) * 2 + 0]
#define LY(l) lm[(l) * 2 + 1]

/* ---- simple.py -------------------------------------------------------------------------- */
static void observe_simple(const mpe_desc *d, const real *pv, const real *lm, real *obs, real *rew) {
    const int A = d->n_agents, L = d->n_landmarks;
    real *o = obs;
    for (int i = 0; i < A; ++i) {
        *o++ = VX(i); *o++ = VY(i);                                         /* simple.py:50 */
        for (int l = 0; l < L; ++l) { *o++ = LX(l) - PX(i); *o++ = LY(l) - PY(i); }   /* :48-49 */
        real dx = PX(i) - LX(0), dy = PY(i) - LY(0);
        rew[i] = -(dx * dx + dy * dy);                                      /* :41-43 */
    }
}

/* ---- simple_spread.py ------------------------------------------------------------------- */
static void observe_spread(const mpe_desc *d, const real *pv, const real *lm, const real *comm,
                           real *obs, real *rew, real *info) {
    const int A = d->n_agents, L = d->n_landmarks, C = d->dim_c;
    real *o = obs;
    /* the landmark term is identical for every agent (simple_spread.py:75-77) */
    real min_sum_rew = 0, min_dists = 0;
    int occupied = 0;
    for (int l = 0; l < L; ++l) {
        real m = dist2d(PX(0), PY(0), LX(l), LY(l));
        for (int a = 1; a < A; ++a) {
            real dd = dist2d(PX(a), PY(a), LX(l), LY(l));
            if (dd < m) m = dd;
        }
        min_sum_rew -= m;                                                   /* :77 / :55 */
        min_dists += m;                                                     /* :54 */
        if (m < R(0.1)) occupied += 1;                                      /* :56-57 */
    }
    for (int i = 0; i < A; ++i) {
        *o++ = VX(i); *o++ = VY(i); *o++ = PX(i); *o++ = PY(i);             /* :100 */
        for (int l = 0; l < L; ++l) { *o++ = LX(l) - PX(i); *o++ = LY(l) - PY(i); }   /* :87-88 */
        for (int j = 0; j < A; ++j) if (j != i) { *o++ = PX(j) - PX(i); *o++ = PY(j) - PY(i); }   /* :99 */
        for (int j = 0; j < A; ++j) if (j != i) for (int q = 0; q < C; ++q) *o++ = comm[j * C + q];  /* :98 */
        real r = min_sum_rew;
        int collisions = 0;
        if (d->agent_collide[i])                                            /* :78-81 (includes a == i) */
            for (int a = 0; a < A; ++a)
                if (is_collision(PX(a), PY(a), R(d->agent_size[a]), PX(i), PY(i), R(d->agent_size[i]))) {
                    r -= R(1.0);
                    collisions += 1;
                }
        rew[i] = r;
        if (info) { info[i * 4 + 0] = r; info[i * 4 + 1] = (real)collisions; info[i * 4 + 2] = min_dists;
                    info[i * 4 + 3] = (real)occupied; }                     /* :47-63 */
    }
}

/* ---- simple_tag.py ---------------------------------------------------------------------- */
static void observe_tag(const mpe_desc *d, const real *pv, const real *lm, real *obs, real *rew, real *info) {
    const int A = d->n_agents, L = d->n_landmarks;
    real *o = obs;
    for (int i = 0; i < A; ++i) {
        *o++ = VX(i); *o++ = VY(i); *o++ = PX(i); *o++ = PY(i);             /* simple_tag.py:147 */
        for (int l = 0; l < L; ++l) { *o++ = LX(l) - PX(i); *o++ = LY(l) - PY(i); }   /* :133-136 */
        for (int j = 0; j < A; ++j) if (j != i) { *o++ = PX(j) - PX(i); *o++ = PY(j) - PY(i); }   /* :144 */
        for (int j = 0; j < A; ++j) if (j != i && !d->agent_adversary[j]) { *o++ = VX(j); *o++ = VY(j); }  /* :145-146 */
        real r = 0;
        int coll = 0;
        if (d->agent_adversary[i]) {                                        /* adversary_reward :115-129 */
            if (d->agent_collide[i])
                for (int g = 0; g < A; ++g) if (!d->agent_adversary[g])
                    for (int a = 0; a < A; ++a) if (d->agent_adversary[a])
                        if (is_collision(PX(g), PY(g), R(d->agent_size[g]), PX(a), PY(a), R(d->agent_size[a])))
                            r += R(10.0);
            for (int g = 0; g < A; ++g) if (!d->agent_adversary[g])         /* benchmark_data :57-66 */
                if (is_collision(PX(g), PY(g), R(d->agent_size[g]), PX(i), PY(i), R(d->agent_size[i]))) coll += 1;
        } else {                                                            /* agent_reward :89-113 */
            if (d->agent_collide[i])
                for (int a = 0; a < A; ++a) if (d->agent_adversary[a])
                    if (is_collision(PX(a), PY(a), R(d->agent_size[a]), PX(i), PY(i), R(d->agent_size[i])))
                        r -= R(10.0);
            r -= bound_pen(r_abs(PX(i)));                                   /* :109-111 */
            r -= bound_pen(r_abs(PY(i)));
        }
        rew[i] = r;
        if (info) info[i] = (real)coll;
    }
}

/* ---- simple_world_comm.py --------------------------------------------------------------- */
static void observe_world_comm(const mpe_desc *d, const real *pv, const real *lm, const real *comm,
                               real *obs, real *rew, real *info) {
    const int A = d->n_agents, L = d->n_landmarks, C = d->dim_c;
    const int food0 = d->n_obstacles, forest0 = d->n_obstacles + d->n_food;
    int inf[MAXA][2];
    for (int i = 0; i < A; ++i)
        for (int f = 0; f < 2; ++f)                                         /* :231-239,251-252 */
            inf[i][f] = is_collision(PX(i), PY(i), R(d->agent_size[i]), LX(forest0 + f), LY(forest0 + f),
                                     R(d->landmark_size[forest0 + f]));
    int leader = 0;
    for (int i = 0; i < A; ++i) if (d->agent_leader[i]) leader = i;
    real *o = obs;
    for (int i = 0; i < A; ++i) {
        *o++ = VX(i); *o++ = VY(i); *o++ = PX(i); *o++ = PY(i);             /* :281-287 */
        for (int l = 0; l < L; ++l) { *o++ = LX(l) - PX(i); *o++ = LY(l) - PY(i); }   /* :226-229 */
        int vis[MAXA];
        for (int j = 0; j < A; ++j)                                         /* :253 */
            vis[j] = (inf[i][0] && inf[j][0]) || (inf[i][1] && inf[j][1]) ||
                     (!inf[i][0] && !inf[j][0] && !inf[i][1] && !inf[j][1]) || d->agent_leader[i];
        for (int j = 0; j < A; ++j) if (j != i) {                           /* other_pos :254,258 */
            *o++ = vis[j] ? PX(j) - PX(i) : R(0.0);
            *o++ = vis[j] ? PY(j) - PY(i) : R(0.0);
        }
        if (d->agent_adversary[i]) {
            for (int j = 0; j < A; ++j) if (j != i && !d->agent_adversary[j]) {   /* other_vel :255-256,259-260 */
                *o++ = vis[j] ? VX(j) : R(0.0);
                *o++ = vis[j] ? VY(j) : R(0.0);
            }
            *o++ = inf[i][0] ? R(1.0) : R(-1.0);                            /* in_forest */
            *o++ = inf[i][1] ? R(1.0) : R(-1.0);
            for (int q = 0; q < C; ++q) *o++ = comm[leader * C + q];        /* :279 comm = [agents[0].state.c] */
        } else {
            *o++ = inf[i][0] ? R(1.0) : R(-1.0);                            /* :287: in_forest before other_vel */
            *o++ = inf[i][1] ? R(1.0) : R(-1.0);
            for (int j = 0; j < A; ++j) if (j != i && !d->agent_adversary[j]) {
                *o++ = vis[j] ? VX(j) : R(0.0);
                *o++ = vis[j] ? VY(j) : R(0.0);
            }
        }
        real r = 0;
        int coll = 0;
        if (d->agent_adversary[i]) {                                        /* adversary_reward :185-198 */
            real m = 0;
            int first = 1;
            for (int g = 0; g < A; ++g) if (!d->agent_adversary[g]) {
                real dd = dist2d(PX(g), PY(g), PX(i), PY(i));
                if (first || dd < m) m = dd;
                first = 0;
            }
            r -= R(0.1) * m;                                                /* :192 */
            if (d->agent_collide[i])
                for (int g = 0; g < A; ++g) if (!d->agent_adversary[g])
                    for (int a = 0; a < A; ++a) if (d->agent_adversary[a])
                        if (is_collision(PX(g), PY(g), R(d->agent_size[g]), PX(a), PY(a), R(d->agent_size[a])))
                            r += R(5.0);                                    /* :193-197 */
            for (int g = 0; g < A; ++g) if (!d->agent_adversary[g])         /* benchmark_data :115-123 */
                if (is_collision(PX(g), PY(g), R(d->agent_size[g]), PX(i), PY(i), R(d->agent_size[i]))) coll += 1;
        } else {                                                            /* agent_reward :155-183 */
            if (d->agent_collide[i])
                for (int a = 0; a < A; ++a) if (d->agent_adversary[a])
                    if (is_collision(PX(a), PY(a), R(d->agent_size[a]), PX(i), PY(i), R(d->agent_size[i])))
                        r -= R(5.0);                                        /* :163-166 */
            r -= R(2.0) * bound_pen(r_abs(PX(i)));                          /* :176-178 */
            r -= R(2.0) * bound_pen(r_abs(PY(i)));
            real m = 0;
            for (int f = 0; f < d->n_food; ++f) {
                if (is_collision(PX(i), PY(i), R(d->agent_size[i]), LX(food0 + f), LY(food0 + f),
                                 R(d->landmark_size[food0 + f])))
                    r += R(2.0);                                            /* :179-181 */
            }
            for (int f = 0; f < d->n_food; ++f) {
                real dd = dist2d(LX(food0 + f), LY(food0 + f), PX(i), PY(i));
                if (f == 0 || dd < m) m = dd;
            }
            r += R(0.05) * m;                                               /* :182 */
        }
        rew[i] = r;
        if (info) info[i] = (real)coll;
    }
}

/* ---- simple_adversary.py ----------------------------------------------------------------- */
/* goal[0] = index of the goal landmark (np.random.choice(world.landmarks), :44) */
static void observe_adversary(const mpe_desc *d, const real *pv, const real *lm, const int32_t *goal,
                              real *obs, real *rew, real *info) {
    const int A = d->n_agents, L = d->n_landmarks, g = goal[0], I = L + 1;
    real *o = obs;
    /* agent_reward terms (:76-105): shaped, sum over adversaries / min over good agents */
    real adv_sum = 0, good_min = 0;
    int first = 1;
    for (int a = 0; a < A; ++a) {
        real dd = dist2d(PX(a), PY(a), LX(g), LY(g));
        if (d->agent_adversary[a]) adv_sum += dd;                          /* :83 */
        else { if (first || dd < good_min) good_min = dd; first = 0; }      /* :93-94 */
    }
    for (int i = 0; i < A; ++i) {
        if (!d->agent_adversary[i]) { *o++ = LX(g) - PX(i); *o++ = LY(g) - PY(i); }          /* :136 */
        for (int l = 0; l < L; ++l) { *o++ = LX(l) - PX(i); *o++ = LY(l) - PY(i); }           /* :123-125 */
        for (int j = 0; j < A; ++j) if (j != i) { *o++ = PX(j) - PX(i); *o++ = PY(j) - PY(i); }   /* :131-133 */
        real dx = PX(i) - LX(g), dy = PY(i) - LY(g);
        real d2 = dx * dx + dy * dy;
        if (d->agent_adversary[i]) rew[i] = -d2;                            /* adversary_reward :107-118 */
        else rew[i] = -good_min + adv_sum;                                  /* :105 pos_rew + adv_rew */
        if (info) {                                                         /* benchmark_data :61-70 */
            for (int q = 0; q < I; ++q) info[i * I + q] = 0;
            if (d->agent_adversary[i]) info[i * I] = d2;
            else {
                for (int l = 0; l < L; ++l) { real ex = PX(i) - LX(l), ey = PY(i) - LY(l); info[i * I + l] = ex * ex + ey * ey; }
                info[i * I + L] = d2;
            }
        }
    }
}

/* ---- simple_push.py ---------------------------------------------------------------------- */
static void observe_push(const mpe_desc *d, const real *pv, const real *lm, const int32_t *goal, real *obs, real *rew) {
    const int A = d->n_agents, L = d->n_landmarks, g = goal[0];
    real *o = obs;
    real good_min = 0;
    int first = 1;
    for (int a = 0; a < A; ++a) if (!d->agent_adversary[a]) {              /* adversary_reward :68-69 */
        real dd = dist2d(PX(a), PY(a), LX(g), LY(g));
        if (first || dd < good_min) good_min = dd;
        first = 0;
    }
    for (int i = 0; i < A; ++i) {
        *o++ = VX(i); *o++ = VY(i);
        if (!d->agent_adversary[i]) {                                       /* :93-94 */
            *o++ = LX(g) - PX(i); *o++ = LY(g) - PY(i);
            /* agent.color = [0.25]*3 with color[goal.index + 1] += 0.5 (:47-53) */
            for (int c = 0; c < 3; ++c) *o++ = (c == g + 1) ? R(0.25) + R(0.5) : R(0.25);
        }
        for (int l = 0; l < L; ++l) { *o++ = LX(l) - PX(i); *o++ = LY(l) - PY(i); }
        if (!d->agent_adversary[i])                                         /* entity_color: [0.1]*3, color[i+1] += 0.8 (:34-37) */
            for (int l = 0; l < L; ++l) for (int c = 0; c < 3; ++c) *o++ = (c == l + 1) ? R(0.1) + R(0.8) : R(0.1);
        for (int j = 0; j < A; ++j) if (j != i) { *o++ = PX(j) - PX(i); *o++ = PY(j) - PY(i); }
        if (d->agent_adversary[i])                                          /* :66-74 */
            rew[i] = good_min - dist2d(LX(g), LY(g), PX(i), PY(i));
        else                                                                /* :62-64 */
            rew[i] = -dist2d(PX(i), PY(i), LX(g), LY(g));
    }
}

/* ---- simple_speaker_listener.py ---------------------------------------------------------- */
/* goal[0] = agents[0].goal_b; landmark colours [0.65,0.15,0.15],[0.15,0.65,0.15],[0.15,0.15,0.65] (:44-46) */
static void observe_speaker_listener(const mpe_desc *d, const real *pv, const real *lm, const real *comm,
                                     const int32_t *goal, real *obs, real *rew) {
    const int L = d->n_landmarks, C = d->dim_c, g = goal[0];
    real *o = obs;
    for (int c = 0; c < 3; ++c) *o++ = (c == g) ? R(0.65) : R(0.15);        /* speaker: goal_color :70-72,87-88 */
    *o++ = VX(1); *o++ = VY(1);                                             /* listener :90-92 */
    for (int l = 0; l < L; ++l) { *o++ = LX(l) - PX(1); *o++ = LY(l) - PY(1); }
    for (int q = 0; q < C; ++q) *o++ = comm[0 * C + q];
    real dx = PX(1) - LX(g), dy = PY(1) - LY(g);                            /* reward :63-67: a = agents[0] */
    rew[0] = rew[1] = -(dx * dx + dy * dy);
}

/* ---- simple_reference.py ----------------------------------------------------------------- */
/* goal[i] = agents[i].goal_b; agents[i].goal_a = the other agent; landmark colours (:37-39) */
static void observe_reference(const mpe_desc *d, const real *pv, const real *lm, const real *comm,
                              const int32_t *goal, real *obs, real *rew) {
    const int A = d->n_agents, L = d->n_landmarks, C = d->dim_c;
    real *o = obs;
    for (int i = 0; i < A; ++i) {
        const int other = 1 - i, g = goal[i];
        *o++ = VX(i); *o++ = VY(i);                                         /* :80 */
        for (int l = 0; l < L; ++l) { *o++ = LX(l) - PX(i); *o++ = LY(l) - PY(i); }
        for (int c = 0; c < 3; ++c) *o++ = (c == g) ? R(0.75) : R(0.25);    /* goal_color[1] = goal_b.color :64-66 */
        for (int q = 0; q < C; ++q) *o++ = comm[other * C + q];             /* :76-79 */
        real dx = PX(other) - LX(g), dy = PY(other) - LY(g);                /* reward :55-59 */
        rew[i] = -(dx * dx + dy * dy);
    }
}

/* ---- simple_crypto.py -------------------------------------------------------------------- */
/* goal[0] = goal landmark, goal[1] = key landmark; colours are one-hot in dim_c (:58-62) */
static void observe_crypto(const mpe_desc *d, const real *comm, const int32_t *goal, real *obs, real *rew, real *info) {
    const int A = d->n_agents, C = d->dim_c, g = goal[0], key = goal[1], I = 2 * C;
    real *o = obs;
    real err[MAXA];
    int spoke[MAXA];
    for (int a = 0; a < A; ++a) {
        real e = 0;
        int nz = 0;
        for (int q = 0; q < C; ++q) {
            real df = comm[a * C + q] - (q == g ? R(1.0) : R(0.0));
            e += df * df;
            if (comm[a * C + q] != 0) nz = 1;
        }
        err[a] = e; spoke[a] = nz;
    }
    for (int i = 0; i < A; ++i) {
        const int speaker = (i == 2);
        if (speaker) {                                                      /* :157-162 */
            for (int q = 0; q < C; ++q) *o++ = (q == g) ? R(1.0) : R(0.0);
            for (int q = 0; q < C; ++q) *o++ = (q == key) ? R(1.0) : R(0.0);
        } else if (!d->agent_adversary[i]) {                                /* listener :163-168 */
            for (int q = 0; q < C; ++q) *o++ = (q == key) ? R(1.0) : R(0.0);
            for (int q = 0; q < C; ++q) *o++ = comm[2 * C + q];
        } else {                                                            /* adversary :169-174 */
            for (int q = 0; q < C; ++q) *o++ = comm[2 * C + q];
        }
        real r = 0;
        if (d->agent_adversary[i]) {                                        /* adversary_reward :115-121 */
            if (spoke[i]) r -= err[i];
        } else {                                                            /* agent_reward :94-113 */
            real good_rew = 0, adv_rew = 0;
            for (int a = 0; a < A; ++a) if (!d->agent_adversary[a] && a != 2 && spoke[a]) good_rew -= err[a];
            for (int a = 0; a < A; ++a) if (d->agent_adversary[a] && spoke[a]) adv_rew += err[a];
            r = adv_rew + good_rew;
        }
        rew[i] = r;
        if (info) {                                                         /* benchmark_data :66-67 */
            for (int q = 0; q < C; ++q) info[i * I + q] = comm[i * C + q];
            for (int q = 0; q < C; ++q) info[i * I + C + q] = (q == g) ? R(1.0) : R(0.0);
        }
    }
}

static void observe_world(const mpe_desc *d, const real *pv, const real *lm, const real *comm,
                          const int32_t *goal, real *obs, real *rew, uint8_t *done, real *info,
                          uint32_t flags) {
    const int A = d->n_agents;
    switch (d->scenario) {
    case MPE_SCN_SIMPLE: observe_simple(d, pv, lm, obs, rew); break;
    case MPE_SCN_SPREAD: observe_spread(d, pv, lm, comm, obs, rew, info); break;
    case MPE_SCN_TAG: observe_tag(d, pv, lm, obs, rew, info); break;
    case MPE_SCN_WORLD_COMM: observe_world_comm(d, pv, lm, comm, obs, rew, info); break;
    case MPE_SCN_ADVERSARY: observe_adversary(d, pv, lm, goal, obs, rew, info); break;
    case MPE_SCN_PUSH: observe_push(d, pv, lm, goal, obs, rew); break;
    case MPE_SCN_SPEAKER_LISTENER: observe_speaker_listener(d, pv, lm, comm, goal, obs, rew); break;
    case MPE_SCN_REFERENCE: observe_reference(d, pv, lm, comm, goal, obs, rew); break;
    case MPE_SCN_CRYPTO: observe_crypto(d, comm, goal, obs, rew, info); break;
    default: break;
    }
    /* MultiAgentEnv.step glue (environment.py:95,100-102): done_callback is None -> False;
       shared reward = np.sum(reward_n) for every agent */
    for (int i = 0; i < A; ++i) done[i] = 0;
    if (flags & MPE_FLAG_SHARED_REWARD) {
        real s = 0;
        for (int i = 0; i < A; ++i) s += rew[i];
        for (int i = 0; i < A; ++i) rew[i] = s;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* batch drivers                                                                               */

static int sum_act_dim(const mpe_desc *d) { int s = 0; for (int i = 0; i < d->n_agents; ++i) s += FN(mpe_oracle_act_dim)(d, i); return s; }
static int sum_obs_dim(const mpe_desc *d) { int s = 0; for (int i = 0; i < d->n_agents; ++i) s += FN(mpe_oracle_obs_dim)(d, i); return s; }
static int sum_disc_dim(const mpe_desc *d) { int s = 0; for (int i = 0; i < d->n_agents; ++i) s += (d->agent_movable[i] ? 1 : 0) + (d->agent_silent[i] ? 0 : 1); return s; }

void FN(mpe_oracle_set_action)(const mpe_desc *d, int64_t n, const real *act, uint32_t flags, real *u, real *c) {
    const int sa = (flags & MPE_FLAG_DISCRETE_ACTION_INPUT) ? sum_disc_dim(d) : sum_act_dim(d);
    for (int64_t w = 0; w < n; ++w)
        set_action_world(d, act + w * sa, flags, u + w * d->n_agents * 2, c + w * d->n_agents * d->dim_c);
}

void FN(mpe_oracle_world_step)(const mpe_desc *d, int64_t n, real *pv, const real *lm, real *comm,
                               const real *u, const real *c) {
    const int A = d->n_agents, L = d->n_landmarks, C = d->dim_c;
    for (int64_t w = 0; w < n; ++w)
        world_step_world(d, pv + w * A * 4, lm + w * L * 2, comm + w * A * C, u + w * A * 2, c + w * A * C);
}

void FN(mpe_oracle_observe)(const mpe_desc *d, int64_t n, const real *pv, const real *lm, const real *comm,
                            const int32_t *goal, int n_goal, real *obs, real *rew, uint8_t *done,
                            real *info, uint32_t flags) {
    const int A = d->n_agents, L = d->n_landmarks, C = d->dim_c, so = sum_obs_dim(d);
    const int idim = FN(mpe_oracle_info_dim)(d);
    for (int64_t w = 0; w < n; ++w)
        observe_world(d, pv + w * A * 4, lm + w * L * 2, comm + w * A * C, goal ? goal + w * n_goal : 0,
                      obs + w * so, rew + w * A, done + w * A, info ? info + w * A * idim : 0, flags);
}

/* MultiAgentEnv.step (environment.py:80-104) */
void FN(mpe_oracle_step)(const mpe_desc *d, int64_t n, real *pv, const real *lm, real *comm,
                         const int32_t *goal, int n_goal, const real *act, real *obs, real *rew,
                         uint8_t *done, real *info, uint32_t flags) {
    const int A = d->n_agents, L = d->n_landmarks, C = d->dim_c, so = sum_obs_dim(d);
    const int sa = (flags & MPE_FLAG_DISCRETE_ACTION_INPUT) ? sum_disc_dim(d) : sum_act_dim(d);
    const int idim = FN(mpe_oracle_info_dim)(d);
    for (int64_t w = 0; w < n; ++w) {
        real u[MAXA * 2], c[MAXA * MAXC];
        set_action_world(d, act + w * sa, flags, u, c);
        world_step_world(d, pv + w * A * 4, lm + w * L * 2, comm + w * A * C, u, c);
        observe_world(d, pv + w * A * 4, lm + w * L * 2, comm + w * A * C, goal ? goal + w * n_goal : 0,
                      obs + w * so, rew + w * A, done + w * A, info ? info + w * A * idim : 0, flags);
    }
}

/* `steps` consecutive MultiAgentEnv.step calls on the same action batch, entirely in C (used by
 * bench.py to time this port on all host threads without returning to Python between steps) */
void FN(mpe_oracle_rollout)(const mpe_desc *d, int64_t n, real *pv, const real *lm, real *comm,
                            const real *act, int steps, real *obs, real *rew, uint8_t *done, uint32_t flags) {
    for (int t = 0; t < steps; ++t)
        FN(mpe_oracle_step)(d, n, pv, lm, comm, 0, 0, act, obs, rew, done, 0, flags);
}
