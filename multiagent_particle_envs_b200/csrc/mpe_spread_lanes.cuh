// mpe_spread_lanes.cuh -- lane-per-AGENT variant of the fused step for simple_spread.
//
// The lane-per-world kernel (mpe_kernels.cu) gives a batch of 65 536 worlds only 2048 warps, 0.6 of a wave on
// 148 SMs: ~3 warps per scheduler, each dragging a ~700-instruction dependent chain.  Here a world is spread
// over G = 2 / 4 / 8 adjacent lanes (one per agent, padded to a power of two), so the same batch yields
// 8192 / 16384 warps with ~3x shorter chains.  Agents exchange positions with __shfl_sync inside their
// G-lane group; the landmark term of the reward is a shuffle min-reduction over the group and the shared
// reward a shuffle gather (this is the "warp-shuffle reductions for min-distance and collision counts"
// formulation).  Every floating-point operation is the same explicit primitive, applied in the same
// order, as in the lane-per-world kernel, so the two kernels are bit-identical (tests compare them).
#pragma once
#include "mpe_scenarios.cuh"

namespace mpe {

template <int N>
struct SpreadLanes {
    using P = Spread<N>;
    static constexpr int G = N <= 2 ? 2 : (N <= 4 ? 4 : 8);   // lanes per world
    static constexpr int WPW = 32 / G;                         // worlds per warp
    static constexpr int OD = P::obs_dim(0);
    static constexpr int AD = 5;
    static constexpr int kActTile = WPW * AD;                  // floats; rows of one agent, contiguous in global memory
    // observation tile of one agent: WPW rows x OD floats + a pad chosen so that the 8-byte row stores of a
    // half-warp (lanes = (world, agent) pairs) fall into distinct banks
    __host__ __device__ static constexpr int obs_tile_units() {
        constexpr int row = OD / 2;                            // 8-byte units per row
        for (int pad = 0; pad < 32; pad += 2) {       // even: tiles stay 16-byte aligned
            const int t = WPW * row + pad;
            bool ok = true;
            for (int half = 0; half < 2 && ok; ++half) {
                bool used[16] = {};
                for (int l = 0; l < 16 && ok; ++l) {
                    const int lane = half * 16 + l, sub = lane % G, wl = lane / G;
                    if (sub >= N) continue;
                    const int b = (sub * t + wl * row) % 16;
                    if (used[b]) ok = false;
                    used[b] = true;
                }
            }
            if (ok) return t;
        }
        return WPW * row;
    }
    static constexpr int kObsTile = obs_tile_units() * 2;      // floats
    static constexpr int kBarFloats = 4;
    static constexpr int kActOff = kBarFloats;
    static constexpr int kObsOff = (kActOff + N * kActTile + 3) & ~3;
    static constexpr int kWarpFloats = (kObsOff + N * kObsTile + 3) & ~3;
    static constexpr int kWarpBytes = kWarpFloats * 4;
};

template <int N>
__global__ void __launch_bounds__(kMaxThreads) spread_lanes_kernel(const __grid_constant__ StepArgs a) {
    using S = SpreadLanes<N>;
    using P = Spread<N>;
    constexpr int G = S::G, WPW = S::WPW, OD = S::OD, AD = S::AD, INFO = P::INFO;
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % G, wl = lane / G;                   // agent index within the world, world within the warp
    const int64_t n = a.n, end = a.begin + a.count;
    const int64_t w0 = a.begin + (static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp) * WPW;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (w0 >= end) return;
    const int rows = (end - w0) < WPW ? static_cast<int>(end - w0) : WPW;
    const bool agent = sub < N, active = agent && wl < rows;
    const int ai = agent ? sub : 0;
    const int64_t wi = w0 + (wl < rows ? wl : 0);
    float *s_warp = smem + warp * S::kWarpFloats;
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_warp);
    const DevDesc &d = a.d;

    // ---- action tiles: one TMA bulk copy per agent, issued before the state loads ----------------------
    uintptr_t bits = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) bits |= reinterpret_cast<uintptr_t>(a.act[i]);
    const bool bulk = (rows == WPW) && ((bits & 15u) == 0);
    if (bulk && lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, N * S::kActTile * 4);
#pragma unroll
        for (int i = 0; i < N; ++i) bulk_g2s(s_warp + S::kActOff + i * S::kActTile, a.act[i] + w0 * AD, S::kActTile * 4, bar);
    }

    // ---- state: own agent, and landmark number `sub` (A == L in this scenario) ---------------------------
    const float4 pv = a.pv[ai * n + wi];
    const float2 lmv = a.lm[ai * n + wi];
    float px = pv.x, py = pv.y, vx = pv.z, vy = pv.w;

    // ---- MultiAgentEnv._set_action (environment.py:144-192) ----------------------------------------------
    float p0, p1, p2, p3, p4;
    if (bulk) {
        __syncwarp();
        mbar_wait(bar, 0);
        const float *row = s_warp + S::kActOff + ai * S::kActTile + wl * AD;
        p0 = row[0]; p1 = row[1]; p2 = row[2]; p3 = row[3]; p4 = row[4];
    } else {
        const float *row = a.act[ai] + wi * AD;
        p0 = row[0]; p1 = row[1]; p2 = row[2]; p3 = row[3]; p4 = row[4];
    }
    if (a.flags & MPE_FLAG_FORCE_DISCRETE_ACTION) {                 // :169-172 (first arg-max)
        int best = 0;
        float bv = p0;
        if (p1 > bv) { bv = p1; best = 1; }
        if (p2 > bv) { bv = p2; best = 2; }
        if (p3 > bv) { bv = p3; best = 3; }
        if (p4 > bv) { bv = p4; best = 4; }
        p1 = best == 1 ? 1.0f : 0.0f; p2 = best == 2 ? 1.0f : 0.0f;
        p3 = best == 3 ? 1.0f : 0.0f; p4 = best == 4 ? 1.0f : 0.0f;
    }
    const float sens = d.a_sens[ai], size = d.a_size[ai];
    float fx = __fmul_rn(p1 - p2, sens), fy = __fmul_rn(p3 - p4, sens);   // :174-181; apply_action_force core.py:134-140

    // ---- World.step: forces from the other agents, in the lane-per-world kernel's accumulation order ------
    float ox[N], oy[N], osz[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        ox[j] = __shfl_sync(kFull, px, j, G);
        oy[j] = __shfl_sync(kFull, py, j, G);
        osz[j] = d.a_size[j];
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
        // pair (min(i,j), max(i,j)) of core.py:143-155 seen from agent i: force = +pair_force(p_i - p_j), an odd function
        const float2 f = pair_force(__fsub_rn(px, ox[j]), __fsub_rn(py, oy[j]), __fadd_rn(size, osz[j]),
                                    d.contact_force, d.contact_margin, d.inv_margin);
        const bool other = (j != sub);                               // the self pair (dist 0 -> NaN) is discarded
        fx = other ? __fadd_rn(fx, f.x) : fx;
        fy = other ? __fadd_rn(fy, f.y) : fy;
    }
    {
        const float4 r = integrate_entity<false>(px, py, vx, vy, fx, fy, d.keep, d.a_dt_over_mass[ai], d.dt, 0.0f);
        px = r.x; py = r.y; vx = r.z; vy = r.w;
    }
    if (active) a.pv[ai * n + wi] = make_float4(px, py, vx, vy);

    // ---- post-step exchange ------------------------------------------------------------------------------
    float lx[N], ly[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        ox[j] = __shfl_sync(kFull, px, j, G);
        oy[j] = __shfl_sync(kFull, py, j, G);
        lx[j] = __shfl_sync(kFull, lmv.x, j, G);
        ly[j] = __shfl_sync(kFull, lmv.y, j, G);
    }

    // ---- reward (simple_spread.py:72-82): min over agents per landmark = shuffle min-reduction ------------
    float base = 0.0f, min_dists = 0.0f;
    int occupied = 0;
#pragma unroll
    for (int l = 0; l < N; ++l) {
        float m = agent ? dist2d(px, py, lx[l], ly[l]) : __int_as_float(0x7f800000);
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) m = fminf(m, __shfl_xor_sync(kFull, m, off, G));
        base -= m;
        min_dists += m;                                               // benchmark_data :54
        occupied += (m < 0.1f) ? 1 : 0;                               // :56-57
    }
    float r = base;
    int coll = 0;
#pragma unroll
    for (int j = 0; j < N; ++j)                                       // :78-81 (includes j == i)
        if (is_collision(ox[j], oy[j], osz[j], px, py, size)) {
            r -= 1.0f;
            coll += 1;
        }
    float rew = r;
    if (a.flags & MPE_FLAG_SHARED_REWARD) {                           // environment.py:100-102, summed in agent order
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < N; ++j) s += __shfl_sync(kFull, r, j, G);
        rew = s;
    }
    if (active) {
        a.rew[ai * n + wi] = rew;
        a.done[ai * n + wi] = 0;
        if (a.info != nullptr) {
            a.info[(ai * INFO + 0) * n + wi] = r;
            a.info[(ai * INFO + 1) * n + wi] = static_cast<float>(coll);
            a.info[(ai * INFO + 2) * n + wi] = min_dists;
            a.info[(ai * INFO + 3) * n + wi] = static_cast<float>(occupied);
        }
    }

    // ---- observation row of this agent (simple_spread.py:84-100) -------------------------------------------
    if (rows == WPW) {
        float *row = s_warp + S::kObsOff + ai * S::kObsTile + wl * OD;
        if (agent) {
            float2 *r2 = reinterpret_cast<float2 *>(row);
            r2[0] = make_float2(vx, vy);
            r2[1] = make_float2(px, py);
#pragma unroll
            for (int l = 0; l < N; ++l) r2[2 + l] = make_float2(lx[l] - px, ly[l] - py);
#pragma unroll
            for (int k = 0; k < N - 1; ++k) {
                const int j = k + (k >= sub ? 1 : 0);
                r2[2 + N + k] = make_float2(pick(ox, j) - px, pick(oy, j) - py);
            }
#pragma unroll
            for (int k = 0; k < N - 1; ++k) r2[2 + N + (N - 1) + k] = make_float2(0.0f, 0.0f);
        }
        __syncwarp();
        // stream the N tiles out in global order: tile t holds WPW consecutive rows of obs_n[t]
        constexpr int kVecPerTile = WPW * OD / 4, kVec = N * kVecPerTile;
#pragma unroll
        for (int q0 = 0; q0 < kVec; q0 += 32) {
            const int q = q0 + lane;
            if (q0 + 32 <= kVec || q < kVec) {
                const int t = q / kVecPerTile, e = q - t * kVecPerTile;
                const float4 v = *reinterpret_cast<const float4 *>(s_warp + S::kObsOff + t * S::kObsTile + 4 * e);
                __stcs(reinterpret_cast<float4 *>(a.obs[t] + w0 * OD) + e, v);
            }
        }
    } else if (active) {   // the batch's last, partial warp
        float *g = a.obs[ai] + wi * OD;
        g[0] = vx; g[1] = vy; g[2] = px; g[3] = py;
#pragma unroll
        for (int l = 0; l < N; ++l) { g[4 + 2 * l] = lx[l] - px; g[5 + 2 * l] = ly[l] - py; }
#pragma unroll
        for (int k = 0; k < N - 1; ++k) {
            const int j = k + (k >= sub ? 1 : 0);
            g[4 + 2 * N + 2 * k] = pick(ox, j) - px;
            g[5 + 2 * N + 2 * k] = pick(oy, j) - py;
        }
#pragma unroll
        for (int k = 0; k < 2 * (N - 1); ++k) g[4 + 2 * N + 2 * (N - 1) + k] = 0.0f;
    }
}

}  // namespace mpe
