// mpe_scenarios.cuh -- the scenario programs: observation / reward / benchmark_data of each
// reference scenario file, written against per-lane world registers.  Entity counts are template
// parameters so every loop unrolls; entity *properties* (size, collide, ...) stay runtime values
// read from the constant bank, i.e. whatever Scenario.make_world() put on the World object.
//
// A program P provides
//   A, L, DIMC, NS (#speakers = non-silent agents, the first NS agents), INFO (floats/agent)
//   obs_dim(i), act_dim(i), movable(i)            constexpr shape functions
//   observe<I>(d, w, out)                         scenario.observation(agent I, world)
//   reward(d, w, rew[A], info[A*INFO])            scenario.reward / benchmark_data for every agent
//   validate(desc)                                host: does the descriptor fit this program?
#pragma once
#include "mpe_common.cuh"

namespace mpe {

template <int A_, int L_, int NC_, int NG_ = 0>
struct WorldRegs {
    float px[A_], py[A_], vx[A_], vy[A_];
    float lx[L_], ly[L_];
    float c[NC_ > 0 ? NC_ : 1];  // comm state of the speakers, [speaker][dim_c]
    int g[NG_ > 0 ? NG_ : 1];    // per-world goal indices
    unsigned aux;                // per-world predicate bits filled by P::prepare (world_comm: forest membership)
};

// programs without shared per-world predicates inherit the empty prepare()
struct ProgramBase {
    template <class W>
    __device__ __forceinline__ static void prepare(const DevDesc &, W &) {}
    // true: also build the 80-register variant of the fused step (only for programs that fit without spilling)
    static constexpr bool kLowRegVariant = false;
};

// v[idx] for a per-lane index without dynamic register indexing (select chain, N <= 8)
template <int N>
__device__ __forceinline__ float pick(const float (&v)[N], int idx) {
    float r = v[0];
#pragma unroll
    for (int k = 1; k < N; ++k) r = (idx == k) ? v[k] : r;
    return r;
}


// does the descriptor carry exactly the structural flags a program was compiled for?
template <class P>
static bool structure_matches(const mpe_desc &d) {
    for (int i = 0; i < P::A; ++i) {
        if ((d.agent_collide[i] != 0) != P::agent_collides(i)) return false;
        if ((d.agent_movable[i] != 0) != P::movable(i)) return false;
        if ((d.agent_max_speed[i] >= 0) != P::kSpeedLimit) return false;
    }
    for (int l = 0; l < P::L; ++l)
        if ((d.landmark_collide[l] != 0) != P::landmark_collides(l)) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------
// simple.py : 1 agent, 1 landmark, nothing collides
template <int A_, int L_>
struct Simple : ProgramBase {
    static constexpr int A = A_, L = L_, DIMC = 0, NS = 0, INFO = 0, G = 0;
    static constexpr int kScenario = MPE_SCN_SIMPLE;
    using W = WorldRegs<A, L, 0>;
    __host__ __device__ static constexpr int obs_dim(int) { return 2 + 2 * L; }   // simple.py:45-50
    __host__ __device__ static constexpr int act_dim(int) { return 5; }
    __host__ __device__ static constexpr bool movable(int) { return true; }
    // structural flags fixed by simple.py:6-22 (checked against the descriptor in validate())
    __host__ __device__ static constexpr bool agent_collides(int) { return false; }
    __host__ __device__ static constexpr bool landmark_collides(int) { return false; }
    static constexpr bool kSpeedLimit = false;

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &, const W &w, Wr &o) {
        o.put2(w.vx[I], w.vy[I]);                                                    // :50
#pragma unroll
        for (int l = 0; l < L; ++l) o.put2(sub2(make_float2(w.lx[l], w.ly[l]), make_float2(w.px[I], w.py[I])));    // :48-49
    }
    __device__ __forceinline__ static void reward(const DevDesc &, const W &w, float (&rew)[A], float *) {
#pragma unroll
        for (int i = 0; i < A; ++i) {
            const float dx = w.px[i] - w.lx[0], dy = w.py[i] - w.ly[0];
            rew[i] = -__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));                                           // :41-43
        }
    }
    static bool validate(const mpe_desc &d) {
        return d.n_agents == A && d.n_landmarks == L && d.dim_c == 0 && structure_matches<Simple>(d);
    }
};

// ---------------------------------------------------------------------------------------------
// simple_spread.py : N agents, N landmarks, cooperative
template <int N_>
struct Spread : ProgramBase {
    static constexpr int A = N_, L = N_, DIMC = 2, NS = 0, INFO = 4, G = 0;
    static constexpr int kScenario = MPE_SCN_SPREAD;
    static constexpr bool kLowRegVariant = (N_ == 4);
    using W = WorldRegs<A, L, 0>;
    __host__ __device__ static constexpr int obs_dim(int) { return 4 + 2 * L + 2 * (A - 1) + DIMC * (A - 1); }
    __host__ __device__ static constexpr int act_dim(int) { return 5; }
    __host__ __device__ static constexpr bool movable(int) { return true; }
    // structural flags fixed by simple_spread.py:15-26 (checked against the descriptor in validate())
    __host__ __device__ static constexpr bool agent_collides(int) { return true; }
    __host__ __device__ static constexpr bool landmark_collides(int) { return false; }
    static constexpr bool kSpeedLimit = false;

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &, const W &w, Wr &o) {
        o.put2(w.vx[I], w.vy[I]);                                                    // simple_spread.py:100
        o.put2(w.px[I], w.py[I]);
#pragma unroll
        for (int l = 0; l < L; ++l) o.put2(sub2(make_float2(w.lx[l], w.ly[l]), make_float2(w.px[I], w.py[I])));    // :87-88
#pragma unroll
        for (int j = 0; j < A; ++j)
            if (j != I) o.put2(sub2(make_float2(w.px[j], w.py[j]), make_float2(w.px[I], w.py[I])));                // :99
#pragma unroll
        for (int j = 0; j < (A - 1) * DIMC; ++j) o.put(0.0f);                        // :98 (all agents silent)
    }
    __device__ __forceinline__ static void reward(const DevDesc &d, const W &w, float (&rew)[A], float *info) {
        float base = 0.0f, min_dists = 0.0f;
        int occupied = 0;
#pragma unroll
        for (int l = 0; l < L; ++l) {                                                // :75-77
            float m = dist2d(w.px[0], w.py[0], w.lx[l], w.ly[l]);
#pragma unroll
            for (int a = 1; a < A; ++a) m = fminf(m, dist2d(w.px[a], w.py[a], w.lx[l], w.ly[l]));
            base -= m;
            min_dists += m;                                                          // :54
            occupied += (m < 0.1f) ? 1 : 0;                                          // :56-57
        }
        // is_collision is symmetric in its arguments bit for bit: evaluate each unordered pair once
        bool hit[A][A];
#pragma unroll
        for (int i = 0; i < A; ++i)
#pragma unroll
            for (int a = i; a < A; ++a)
                hit[i][a] = hit[a][i] = is_collision(w.px[a], w.py[a], d.a_size[a], w.px[i], w.py[i], d.a_size[i]);
#pragma unroll
        for (int i = 0; i < A; ++i) {
            float r = base;
            int coll = 0;
            if (agent_collides(i)) {                                                 // :78-81, a == i included
#pragma unroll
                for (int a = 0; a < A; ++a)
                    if (hit[i][a]) {
                        r -= 1.0f;
                        coll += 1;
                    }
            }
            rew[i] = r;
            if (info) {                                                              // benchmark_data :47-63
                info[i * INFO + 0] = r;
                info[i * INFO + 1] = static_cast<float>(coll);
                info[i * INFO + 2] = min_dists;
                info[i * INFO + 3] = static_cast<float>(occupied);
            }
        }
    }
    static bool validate(const mpe_desc &d) {
        if (d.n_agents != A || d.n_landmarks != L || d.dim_c != DIMC) return false;
        for (int i = 0; i < A; ++i)
            if (!d.agent_movable[i] || !d.agent_silent[i]) return false;
        return structure_matches<Spread>(d);
    }
};

// ---------------------------------------------------------------------------------------------
// simple_tag.py : NADV adversaries (first), NGOOD prey, L obstacles
template <int NADV_, int NGOOD_, int L_>
struct Tag : ProgramBase {
    static constexpr int NADV = NADV_, NGOOD = NGOOD_;
    static constexpr int A = NADV + NGOOD, L = L_, DIMC = 2, NS = 0, INFO = 1, G = 0;
    static constexpr int kScenario = MPE_SCN_TAG;
    static constexpr bool kLowRegVariant = (A >= 3 && A <= 6);   // 3+1 and 4+2 fit 80 registers; 6+2 spills
    using W = WorldRegs<A, L, 0>;
    __host__ __device__ static constexpr bool adversary(int i) { return i < NADV; }
    __host__ __device__ static constexpr int obs_dim(int i) {                       // simple_tag.py:131-147
        return 4 + 2 * L + 2 * (A - 1) + 2 * (NGOOD - (adversary(i) ? 0 : 1));
    }
    __host__ __device__ static constexpr int act_dim(int) { return 5; }
    __host__ __device__ static constexpr bool movable(int) { return true; }
    // structural flags fixed by simple_tag.py:16-33 (checked against the descriptor in validate())
    __host__ __device__ static constexpr bool agent_collides(int) { return true; }
    __host__ __device__ static constexpr bool landmark_collides(int) { return true; }
    static constexpr bool kSpeedLimit = true;

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &, const W &w, Wr &o) {
        o.put2(w.vx[I], w.vy[I]);                                                    // :147
        o.put2(w.px[I], w.py[I]);
#pragma unroll
        for (int l = 0; l < L; ++l) o.put2(sub2(make_float2(w.lx[l], w.ly[l]), make_float2(w.px[I], w.py[I])));    // :133-136
#pragma unroll
        for (int j = 0; j < A; ++j)
            if (j != I) o.put2(sub2(make_float2(w.px[j], w.py[j]), make_float2(w.px[I], w.py[I])));                // :144
#pragma unroll
        for (int j = NADV; j < A; ++j)
            if (j != I) o.put2(w.vx[j], w.vy[j]);                                    // :145-146
    }
    __device__ __forceinline__ static void reward(const DevDesc &d, const W &w, float (&rew)[A], float *info) {
        // every (good, adversary) contact flag once; both reward branches are sums over them
        bool hit[NGOOD][NADV];
#pragma unroll
        for (int g = 0; g < NGOOD; ++g)
#pragma unroll
            for (int a = 0; a < NADV; ++a)
                hit[g][a] = is_collision(w.px[NADV + g], w.py[NADV + g], d.a_size[NADV + g],
                                         w.px[a], w.py[a], d.a_size[a]);
#pragma unroll
        for (int i = 0; i < A; ++i) {
            float r = 0.0f;
            int coll = 0;
            if (adversary(i)) {                                                      // adversary_reward :115-129
#pragma unroll
                for (int g = 0; g < NGOOD; ++g) {
#pragma unroll
                    for (int a = 0; a < NADV; ++a)
                        if (agent_collides(i) && hit[g][a]) r += 10.0f;
                    coll += hit[g][i < NADV ? i : 0] ? 1 : 0;                        // benchmark_data :57-66
                }
            } else {                                                                 // agent_reward :89-113
#pragma unroll
                for (int a = 0; a < NADV; ++a)
                    if (agent_collides(i) && hit[i >= NADV ? i - NADV : 0][a]) r -= 10.0f;
                r -= bound_pen(fabsf(w.px[i]));                                      // :109-111
                r -= bound_pen(fabsf(w.py[i]));
            }
            rew[i] = r;
            if (info) info[i] = static_cast<float>(coll);
        }
    }
    static bool validate(const mpe_desc &d) {
        if (d.n_agents != A || d.n_landmarks != L || d.dim_c != DIMC || d.n_adversaries != NADV) return false;
        for (int i = 0; i < A; ++i)
            if (!d.agent_movable[i] || !d.agent_silent[i] || (d.agent_adversary[i] != 0) != adversary(i)) return false;
        return structure_matches<Tag>(d);
    }
};

// ---------------------------------------------------------------------------------------------
// simple_world_comm.py : NADV adversaries (agent 0 = leader, the only speaker), NGOOD prey,
// landmarks = NOBST obstacles ++ NFOOD food ++ 2 forests
template <int NADV_, int NGOOD_, int NOBST_, int NFOOD_>
struct WorldComm : ProgramBase {
    static constexpr int NADV = NADV_, NGOOD = NGOOD_, NOBST = NOBST_, NFOOD = NFOOD_, NFOREST = 2;
    static constexpr int A = NADV + NGOOD, L = NOBST + NFOOD + NFOREST, DIMC = 4, NS = 1, INFO = 1, G = 0;
    static constexpr int FOOD0 = NOBST, FOREST0 = NOBST + NFOOD;
    static constexpr int kScenario = MPE_SCN_WORLD_COMM;
    using W = WorldRegs<A, L, NS * DIMC>;
    __host__ __device__ static constexpr bool adversary(int i) { return i < NADV; }
    __host__ __device__ static constexpr int obs_dim(int i) {                       // simple_world_comm.py:281-287
        return 4 + 2 * L + 2 * (A - 1) + 2 * (NGOOD - (adversary(i) ? 0 : 1)) + 2 + (adversary(i) ? DIMC : 0);
    }
    __host__ __device__ static constexpr int act_dim(int i) { return i == 0 ? 5 + DIMC : 5; }
    __host__ __device__ static constexpr bool movable(int) { return true; }
    // structural flags fixed by simple_world_comm.py:19-50 (checked against the descriptor in validate())
    __host__ __device__ static constexpr bool agent_collides(int) { return true; }
    __host__ __device__ static constexpr bool landmark_collides(int l) { return l < NOBST; }
    static constexpr bool kSpeedLimit = true;

    // forest membership of every agent, once per world: bit 2*i + f = is_collision(agent i, forest f)
    // (:231-239, 251-252).  The reference re-evaluates these 2*A predicates inside every agent's observation
    // (A*A*2 + 2*A evaluations per world); they only depend on the post-step state, so 2*A suffice.
    __device__ __forceinline__ static void prepare(const DevDesc &d, W &w) {
        unsigned bits = 0;
#pragma unroll
        for (int i = 0; i < A; ++i)
#pragma unroll
            for (int f = 0; f < NFOREST; ++f)
                if (is_collision(w.px[i], w.py[i], d.a_size[i], w.lx[FOREST0 + f], w.ly[FOREST0 + f], d.l_size[FOREST0 + f]))
                    bits |= 1u << (2 * i + f);
        w.aux = bits;
    }
    __device__ __forceinline__ static bool in_forest(const DevDesc &, const W &w, int i, int f) {
        return (w.aux >> (2 * i + f)) & 1u;
    }

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &d, const W &w, Wr &o) {
        o.put2(w.vx[I], w.vy[I]);
        o.put2(w.px[I], w.py[I]);
#pragma unroll
        for (int l = 0; l < L; ++l) o.put2(sub2(make_float2(w.lx[l], w.ly[l]), make_float2(w.px[I], w.py[I])));    // :226-229
        const bool f0 = in_forest(d, w, I, 0), f1 = in_forest(d, w, I, 1);
        bool vis[A];
#pragma unroll
        for (int j = 0; j < A; ++j) {                                                // :253 (leader sees all)
            const bool g0 = in_forest(d, w, j, 0), g1 = in_forest(d, w, j, 1);
            vis[j] = (f0 && g0) || (f1 && g1) || (!f0 && !g0 && !f1 && !g1) || (I == 0);
        }
#pragma unroll
        for (int j = 0; j < A; ++j)
            if (j != I) {
                const float2 r = sub2(make_float2(w.px[j], w.py[j]), make_float2(w.px[I], w.py[I]));
                o.put2(vis[j] ? r.x : 0.0f, vis[j] ? r.y : 0.0f);
            }
        if (adversary(I)) {
#pragma unroll
            for (int j = NADV; j < A; ++j)
                if (j != I) o.put2(vis[j] ? w.vx[j] : 0.0f, vis[j] ? w.vy[j] : 0.0f);
            o.put2(f0 ? 1.0f : -1.0f, f1 ? 1.0f : -1.0f);
#pragma unroll
            for (int q = 0; q < DIMC; ++q) o.put(w.c[q]);                            // :279
        } else {
            o.put2(f0 ? 1.0f : -1.0f, f1 ? 1.0f : -1.0f);                            // :287
#pragma unroll
            for (int j = NADV; j < A; ++j)
                if (j != I) o.put2(vis[j] ? w.vx[j] : 0.0f, vis[j] ? w.vy[j] : 0.0f);
        }
    }
    __device__ __forceinline__ static void reward(const DevDesc &d, const W &w, float (&rew)[A], float *info) {
        bool hit[NGOOD][NADV];
#pragma unroll
        for (int g = 0; g < NGOOD; ++g)
#pragma unroll
            for (int a = 0; a < NADV; ++a)
                hit[g][a] = is_collision(w.px[NADV + g], w.py[NADV + g], d.a_size[NADV + g],
                                         w.px[a], w.py[a], d.a_size[a]);
#pragma unroll
        for (int i = 0; i < A; ++i) {
            float r = 0.0f;
            int coll = 0;
            if (adversary(i)) {                                                      // adversary_reward :185-198
                float m = dist2d(w.px[NADV], w.py[NADV], w.px[i], w.py[i]);
#pragma unroll
                for (int g = 1; g < NGOOD; ++g) {
                    const float dd = dist2d(w.px[NADV + g], w.py[NADV + g], w.px[i], w.py[i]);
                    m = dd < m ? dd : m;
                }
                r -= __fmul_rn(0.1f, m);                                                       // :192
#pragma unroll
                for (int g = 0; g < NGOOD; ++g) {
#pragma unroll
                    for (int a = 0; a < NADV; ++a)
                        if (agent_collides(i) && hit[g][a]) r += 5.0f;       // :193-197
                    coll += hit[g][i < NADV ? i : 0] ? 1 : 0;                        // benchmark_data :115-123
                }
            } else {                                                                 // agent_reward :155-183
#pragma unroll
                for (int a = 0; a < NADV; ++a)
                    if (agent_collides(i) && hit[i >= NADV ? i - NADV : 0][a]) r -= 5.0f;
                r -= __fmul_rn(2.0f, bound_pen(fabsf(w.px[i])));                               // :176-178
                r -= __fmul_rn(2.0f, bound_pen(fabsf(w.py[i])));
#pragma unroll
                for (int f = 0; f < NFOOD; ++f)                                      // :179-181
                    if (is_collision(w.px[i], w.py[i], d.a_size[i], w.lx[FOOD0 + f], w.ly[FOOD0 + f],
                                     d.l_size[FOOD0 + f]))
                        r += 2.0f;
                float m = dist2d(w.lx[FOOD0], w.ly[FOOD0], w.px[i], w.py[i]);
#pragma unroll
                for (int f = 1; f < NFOOD; ++f) {
                    const float dd = dist2d(w.lx[FOOD0 + f], w.ly[FOOD0 + f], w.px[i], w.py[i]);
                    m = dd < m ? dd : m;
                }
                r += __fmul_rn(0.05f, m);                                                      // :182
            }
            rew[i] = r;
            if (info) info[i] = static_cast<float>(coll);
        }
    }
    static bool validate(const mpe_desc &d) {
        if (d.n_agents != A || d.n_landmarks != L || d.dim_c != DIMC || d.n_adversaries != NADV) return false;
        if (d.n_obstacles != NOBST || d.n_food != NFOOD || d.n_forests != NFOREST) return false;
        for (int i = 0; i < A; ++i) {
            if (!d.agent_movable[i] || (d.agent_adversary[i] != 0) != adversary(i)) return false;
            if ((d.agent_leader[i] != 0) != (i == 0)) return false;
            if ((d.agent_silent[i] == 0) != (i == 0)) return false;
        }
        return structure_matches<WorldComm>(d);
    }
};

// ---------------------------------------------------------------------------------------------
// simple_adversary.py : NADV adversaries (first), NGOOD good agents, L landmarks, goal = g[0]
template <int NADV_, int NGOOD_, int L_>
struct Adversary : ProgramBase {
    static constexpr int NADV = NADV_, NGOOD = NGOOD_;
    static constexpr int A = NADV + NGOOD, L = L_, DIMC = 2, NS = 0, INFO = L_ + 1, G = 1;
    static constexpr int kScenario = MPE_SCN_ADVERSARY;
    using W = WorldRegs<A, L, 0, G>;
    __host__ __device__ static constexpr bool adversary(int i) { return i < NADV; }
    __host__ __device__ static constexpr int obs_dim(int i) { return (adversary(i) ? 0 : 2) + 2 * L + 2 * (A - 1); }
    __host__ __device__ static constexpr int act_dim(int) { return 5; }
    __host__ __device__ static constexpr bool movable(int) { return true; }
    __host__ __device__ static constexpr bool agent_collides(int) { return false; }    // simple_adversary.py:24
    __host__ __device__ static constexpr bool landmark_collides(int) { return false; }
    static constexpr bool kSpeedLimit = false;

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &, const W &w, Wr &o) {
        if (!adversary(I)) o.put2(sub2(make_float2(pick(w.lx, w.g[0]), pick(w.ly, w.g[0])), make_float2(w.px[I], w.py[I])));   // :136
#pragma unroll
        for (int l = 0; l < L; ++l) o.put2(sub2(make_float2(w.lx[l], w.ly[l]), make_float2(w.px[I], w.py[I])));                  // :123-125
#pragma unroll
        for (int j = 0; j < A; ++j)
            if (j != I) o.put2(sub2(make_float2(w.px[j], w.py[j]), make_float2(w.px[I], w.py[I])));                              // :131-133
    }
    __device__ __forceinline__ static void reward(const DevDesc &, const W &w, float (&rew)[A], float *info) {
        const float gx = pick(w.lx, w.g[0]), gy = pick(w.ly, w.g[0]);
        float adv_sum = 0.0f, good_min = 0.0f;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float dd = dist2d(w.px[a], w.py[a], gx, gy);
            if (adversary(a)) adv_sum += dd;                                                        // :83
            else good_min = (a == NADV) ? dd : (dd < good_min ? dd : good_min);                     // :93-94
        }
#pragma unroll
        for (int i = 0; i < A; ++i) {
            const float dx = __fsub_rn(w.px[i], gx), dy = __fsub_rn(w.py[i], gy);
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            rew[i] = adversary(i) ? -d2 : __fadd_rn(-good_min, adv_sum);                            // :105, :107-118
            if (info) {                                                                             // :61-70
#pragma unroll
                for (int q = 0; q < INFO; ++q) info[i * INFO + q] = 0.0f;
                if (adversary(i)) {
                    info[i * INFO] = d2;
                } else {
#pragma unroll
                    for (int l = 0; l < L; ++l) {
                        const float ex = __fsub_rn(w.px[i], w.lx[l]), ey = __fsub_rn(w.py[i], w.ly[l]);
                        info[i * INFO + l] = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    }
                    info[i * INFO + L] = d2;
                }
            }
        }
    }
    static bool validate(const mpe_desc &d) {
        if (d.n_agents != A || d.n_landmarks != L || d.dim_c != DIMC || d.n_adversaries != NADV) return false;
        for (int i = 0; i < A; ++i)
            if (!d.agent_silent[i] || (d.agent_adversary[i] != 0) != adversary(i)) return false;
        return structure_matches<Adversary>(d);
    }
};

// ---------------------------------------------------------------------------------------------
// simple_push.py : NADV adversaries (first), NGOOD good agents (all collide), L landmarks, goal = g[0]
template <int NADV_, int NGOOD_, int L_>
struct Push : ProgramBase {
    static constexpr int NADV = NADV_, NGOOD = NGOOD_;
    static constexpr int A = NADV + NGOOD, L = L_, DIMC = 2, NS = 0, INFO = 0, G = 1;
    static constexpr int kScenario = MPE_SCN_PUSH;
    using W = WorldRegs<A, L, 0, G>;
    __host__ __device__ static constexpr bool adversary(int i) { return i < NADV; }
    __host__ __device__ static constexpr int obs_dim(int i) {                        // simple_push.py:76-96
        return adversary(i) ? 2 + 2 * L + 2 * (A - 1) : 2 + 2 + 3 + 2 * L + 3 * L + 2 * (A - 1);
    }
    __host__ __device__ static constexpr int act_dim(int) { return 5; }
    __host__ __device__ static constexpr bool movable(int) { return true; }
    __host__ __device__ static constexpr bool agent_collides(int) { return true; }     // simple_push.py:19
    __host__ __device__ static constexpr bool landmark_collides(int) { return false; }
    static constexpr bool kSpeedLimit = false;

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &, const W &w, Wr &o) {
        o.put2(w.vx[I], w.vy[I]);
        if (!adversary(I)) {
            o.put2(sub2(make_float2(pick(w.lx, w.g[0]), pick(w.ly, w.g[0])), make_float2(w.px[I], w.py[I])));       // :93
#pragma unroll
            for (int c = 0; c < 3; ++c) o.put(w.g[0] + 1 == c ? 0.25f + 0.5f : 0.25f); // agent.color (:47-53)
        }
#pragma unroll
        for (int l = 0; l < L; ++l) o.put2(sub2(make_float2(w.lx[l], w.ly[l]), make_float2(w.px[I], w.py[I])));
        if (!adversary(I)) {
#pragma unroll
            for (int l = 0; l < L; ++l)
#pragma unroll
                for (int c = 0; c < 3; ++c) o.put(c == l + 1 ? 0.1f + 0.8f : 0.1f);   // landmark colours (:34-37)
        }
#pragma unroll
        for (int j = 0; j < A; ++j)
            if (j != I) o.put2(sub2(make_float2(w.px[j], w.py[j]), make_float2(w.px[I], w.py[I])));
    }
    __device__ __forceinline__ static void reward(const DevDesc &, const W &w, float (&rew)[A], float *) {
        const float gx = pick(w.lx, w.g[0]), gy = pick(w.ly, w.g[0]);
        float dg[A];
#pragma unroll
        for (int a = 0; a < A; ++a) dg[a] = dist2d(w.px[a], w.py[a], gx, gy);
        float good_min = dg[NADV];
#pragma unroll
        for (int a = NADV + 1; a < A; ++a) good_min = dg[a] < good_min ? dg[a] : good_min;
#pragma unroll
        for (int i = 0; i < A; ++i) rew[i] = adversary(i) ? __fsub_rn(good_min, dg[i]) : -dg[i];   // :62-74
    }
    static bool validate(const mpe_desc &d) {
        if (d.n_agents != A || d.n_landmarks != L || d.dim_c != DIMC || d.n_adversaries != NADV) return false;
        for (int i = 0; i < A; ++i)
            if (!d.agent_silent[i] || (d.agent_adversary[i] != 0) != adversary(i)) return false;
        return structure_matches<Push>(d);
    }
};

// ---------------------------------------------------------------------------------------------
// simple_speaker_listener.py : agent 0 = immovable speaker, agent 1 = silent listener, goal = g[0]
struct SpeakerListener : ProgramBase {
    static constexpr int A = 2, L = 3, DIMC = 3, NS = 1, INFO = 0, G = 1;
    static constexpr int kScenario = MPE_SCN_SPEAKER_LISTENER;
    using W = WorldRegs<A, L, NS * DIMC, G>;
    __host__ __device__ static constexpr int obs_dim(int i) { return i == 0 ? 3 : 2 + 2 * L + DIMC; }
    __host__ __device__ static constexpr bool movable(int i) { return i == 1; }
    __host__ __device__ static constexpr int act_dim(int i) { return i == 0 ? DIMC : 5; }
    __host__ __device__ static constexpr bool agent_collides(int) { return false; }
    __host__ __device__ static constexpr bool landmark_collides(int) { return false; }
    static constexpr bool kSpeedLimit = false;

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &, const W &w, Wr &o) {
        if (I == 0) {                                                                // :70-72, 87-88
#pragma unroll
            for (int c = 0; c < 3; ++c) o.put(w.g[0] == c ? 0.65f : 0.15f);
        } else {                                                                     // :90-92
            o.put2(w.vx[1], w.vy[1]);
#pragma unroll
            for (int l = 0; l < L; ++l) o.put2(sub2(make_float2(w.lx[l], w.ly[l]), make_float2(w.px[1], w.py[1])));
#pragma unroll
            for (int q = 0; q < DIMC; ++q) o.put(w.c[q]);
        }
    }
    __device__ __forceinline__ static void reward(const DevDesc &, const W &w, float (&rew)[A], float *) {
        const float dx = __fsub_rn(w.px[1], pick(w.lx, w.g[0])), dy = __fsub_rn(w.py[1], pick(w.ly, w.g[0]));
        rew[0] = rew[1] = -__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));          // :63-67
    }
    static bool validate(const mpe_desc &d) {
        if (d.n_agents != A || d.n_landmarks != L || d.dim_c != DIMC) return false;
        if (d.agent_silent[0] || !d.agent_silent[1]) return false;
        return structure_matches<SpeakerListener>(d);
    }
};

// ---------------------------------------------------------------------------------------------
// simple_reference.py : 2 agents that move and speak; g[i] = agents[i].goal_b, goal_a = the other agent
struct Reference : ProgramBase {
    static constexpr int A = 2, L = 3, DIMC = 10, NS = 2, INFO = 0, G = 2;
    static constexpr int kScenario = MPE_SCN_REFERENCE;
    using W = WorldRegs<A, L, NS * DIMC, G>;
    __host__ __device__ static constexpr int obs_dim(int) { return 2 + 2 * L + 3 + DIMC * (A - 1); }
    __host__ __device__ static constexpr bool movable(int) { return true; }
    __host__ __device__ static constexpr int act_dim(int) { return 5 + DIMC; }
    __host__ __device__ static constexpr bool agent_collides(int) { return false; }
    __host__ __device__ static constexpr bool landmark_collides(int) { return false; }
    static constexpr bool kSpeedLimit = false;

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &, const W &w, Wr &o) {
        o.put2(w.vx[I], w.vy[I]);                                                    // :80
#pragma unroll
        for (int l = 0; l < L; ++l) o.put2(sub2(make_float2(w.lx[l], w.ly[l]), make_float2(w.px[I], w.py[I])));
#pragma unroll
        for (int c = 0; c < 3; ++c) o.put(w.g[I] == c ? 0.75f : 0.25f);              // goal_b colour :64-66
#pragma unroll
        for (int q = 0; q < DIMC; ++q) o.put(w.c[(1 - I) * DIMC + q]);               // :76-79
    }
    __device__ __forceinline__ static void reward(const DevDesc &, const W &w, float (&rew)[A], float *) {
#pragma unroll
        for (int i = 0; i < A; ++i) {                                                // :55-59
            const float dx = __fsub_rn(w.px[1 - i], pick(w.lx, w.g[i])), dy = __fsub_rn(w.py[1 - i], pick(w.ly, w.g[i]));
            rew[i] = -__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        }
    }
    static bool validate(const mpe_desc &d) {
        if (d.n_agents != A || d.n_landmarks != L || d.dim_c != DIMC) return false;
        if (d.agent_silent[0] || d.agent_silent[1]) return false;
        return structure_matches<Reference>(d);
    }
};

// ---------------------------------------------------------------------------------------------
// simple_crypto.py : Eve (0, adversary), Bob (1), Alice (2, speaker); nobody moves; g[0] = goal, g[1] = key
struct Crypto : ProgramBase {
    static constexpr int A = 3, L = 2, DIMC = 4, NS = 3, INFO = 2 * DIMC, G = 2;
    static constexpr int kScenario = MPE_SCN_CRYPTO;
    using W = WorldRegs<A, L, NS * DIMC, G>;
    __host__ __device__ static constexpr bool adversary(int i) { return i == 0; }
    __host__ __device__ static constexpr int obs_dim(int i) { return i == 0 ? DIMC : 2 * DIMC; }
    __host__ __device__ static constexpr bool movable(int) { return false; }
    __host__ __device__ static constexpr int act_dim(int) { return DIMC; }
    __host__ __device__ static constexpr bool agent_collides(int) { return false; }
    __host__ __device__ static constexpr bool landmark_collides(int) { return false; }
    static constexpr bool kSpeedLimit = false;

    template <int I, class Wr>
    __device__ __forceinline__ static void observe(const DevDesc &, const W &w, Wr &o) {
        if (I == 2) {                                                                // speaker :157-162
#pragma unroll
            for (int q = 0; q < DIMC; ++q) o.put(w.g[0] == q ? 1.0f : 0.0f);
#pragma unroll
            for (int q = 0; q < DIMC; ++q) o.put(w.g[1] == q ? 1.0f : 0.0f);
        } else if (I == 1) {                                                         // listener :163-168
#pragma unroll
            for (int q = 0; q < DIMC; ++q) o.put(w.g[1] == q ? 1.0f : 0.0f);
#pragma unroll
            for (int q = 0; q < DIMC; ++q) o.put(w.c[2 * DIMC + q]);
        } else {                                                                     // adversary :169-174
#pragma unroll
            for (int q = 0; q < DIMC; ++q) o.put(w.c[2 * DIMC + q]);
        }
    }
    __device__ __forceinline__ static void reward(const DevDesc &, const W &w, float (&rew)[A], float *info) {
        float err[A];
        bool spoke[A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            float e = 0.0f;
            bool nz = false;
#pragma unroll
            for (int q = 0; q < DIMC; ++q) {
                const float df = __fsub_rn(w.c[a * DIMC + q], w.g[0] == q ? 1.0f : 0.0f);
                e = __fadd_rn(e, __fmul_rn(df, df));
                nz = nz || (w.c[a * DIMC + q] != 0.0f);
            }
            err[a] = e;
            spoke[a] = nz;
        }
        const float good_rew = spoke[1] ? -err[1] : 0.0f;                             // :98-103
        const float adv_rew = spoke[0] ? err[0] : 0.0f;                              // :104-109
        rew[0] = spoke[0] ? -err[0] : 0.0f;                                          // :115-121
        rew[1] = rew[2] = __fadd_rn(adv_rew, good_rew);                              // :110
        if (info) {                                                                  // :66-67
#pragma unroll
            for (int i = 0; i < A; ++i) {
#pragma unroll
                for (int q = 0; q < DIMC; ++q) {
                    info[i * INFO + q] = w.c[i * DIMC + q];
                    info[i * INFO + DIMC + q] = w.g[0] == q ? 1.0f : 0.0f;
                }
            }
        }
    }
    static bool validate(const mpe_desc &d) {
        if (d.n_agents != A || d.n_landmarks != L || d.dim_c != DIMC || d.n_adversaries != 1) return false;
        for (int i = 0; i < A; ++i)
            if (d.agent_silent[i] || (d.agent_adversary[i] != 0) != adversary(i)) return false;
        return structure_matches<Crypto>(d);
    }
};

}  // namespace mpe
