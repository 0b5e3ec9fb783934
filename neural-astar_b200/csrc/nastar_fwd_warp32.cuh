// nastar_fwd_warp32.cuh — warp-resident forward engine for maps with H <= 32 and W <= 32.
//
// Replaces the T-step loop + backtrack of DifferentiableAstar.forward
// (/root/reference/src/neural_astar/planner/differentiable_astar.py:187-255) for one map per
// warp.  Design (DESIGN.md "warp32 engine"):
//   * lane y owns grid row y: passable / open / closed rows are 32-bit masks in registers,
//     and the lane caches its row's best open cell (order-preserving f key, column).
//   * node selection (:206-209, softmax+argmax == arg-min of (f, flat index), App. A.2) is
//     one REDUX.MIN over the 32 cached row minima + a ballot for the lowest row.
//   * expansion (:228-249) touches <= 8 cells in rows r-1..r+1; those three lanes relax their
//     <= 3 cells from shared memory (g, h, f, parent planes) and fold the new f keys into their
//     cached row minimum — insertions/decreases never need a rescan.
//   * only row r lost its minimum (the selected cell): all 32 lanes rescan that one row
//     (one conflict-free LDS + REDUX.MIN), overlapped with the g2 dependency chain.
//   * planes are staged once per map with 1-D TMA bulk copies (cp.async.bulk + mbarrier) and
//     results leave as coalesced 128-bit stores; the loop never touches HBM.
#pragma once
#include "../../include/nastar_b200.h"
#include "nastar_common.cuh"

namespace nastar {

struct Warp32Smem {
    // byte offsets inside dynamic shared memory for a map of N cells
    int n_pad;
    __host__ __device__ static int npad(int N) { return (N + 3) & ~3; }
    // words per bit array: flat N-bit plane + 1 pad word, and at least one word per lane/row
    __host__ __device__ static int bitwords(int N) {
        const int w = ((N + 31) >> 5) + 1;
        return w < 32 ? 32 : w;
    }
    __host__ __device__ static size_t bytes(int N, bool bwd = false) {
        const int np = npad(N);
        // cost, g, h, f (fp32) + parent (u16) + 2 bit arrays + mbarrier (8 B, 8-B aligned)
        // backward adds the softmax-numerator plane v (fp32, padded to a multiple of 128 cells)
        return size_t(np) * 4 * 4 + size_t(np) * 2 + size_t(2) * bitwords(N) * 4 + 16 +
               (bwd ? size_t((N + 127) & ~127) * 4 + 16 : 0);
    }
};

// Stage one fp32 plane into shared memory. TMA bulk copy when 16-B aligned, LDG/STS otherwise.
// Returns the number of bytes put in flight on the mbarrier (0 if copied synchronously).
__device__ __forceinline__ uint32_t stage_plane(float* dst, const float* src, int N, bool tma_ok, int lane) {
    if (tma_ok) return uint32_t(N) * 4u;
    for (int i = lane; i < N; i += 32) dst[i] = __ldg(src + i);
    return 0u;
}

// Kernel arguments: the forward parameters plus, for the backward replay, the extra fields of
// nastar_bwd_params.  One kernel body serves both so that the replay can never drift from the
// forward's state machine.
struct W32Args {
    nastar_fwd_params f;
    // backward only
    float sqrt_w;
    const int32_t* T_batch;       // device scalar: number of loop iterations of the reference
    const int32_t* t_solve_in;    // forward's t_solve[] (which maps had their goal clamped, App. B)
    const float* grad_hist;
    int64_t grad_stride;
    float* grad_cost;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// kBwd = false: forward search (histories, paths, t_solve, n_steps, optional trace).
// kBwd = true : replays the same search and accumulates the closed-form gradient
//               dL/dcost[p] = sum_t -(1-g_ratio)/sqrt(W) * y_t[p] * (Gh[p] - <Gh, y_t>)   
//               where y_t is the softmax over the open set at step t (differentiable_astar.py:206-209,
//               :55-74) — the only path the reference's autograd keeps alive (:237-243 detach the rest).
template <bool kTrace, bool kBwd>
__global__ void __launch_bounds__(32) astar_warp32_kernel(const W32Args a) {
    const nastar_fwd_params& p = a.f;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x;
    const int H = p.H, W = p.W, N = H * W;
    const int np = Warp32Smem::npad(N);
    const int nwords = (N + 31) >> 5;

    float* sCost = reinterpret_cast<float*>(smem_raw);
    float* sG = sCost + np;
    float* sH = sG + np;
    float* sF = sH + np;
    uint16_t* sPar = reinterpret_cast<uint16_t*>(sF + np);
    uint32_t* sBitsA = reinterpret_cast<uint32_t*>(sPar + np);  // obstacle flat bits, later closed rows
    uint32_t* sBitsB = sBitsA + Warp32Smem::bitwords(N);        // path rows
    uint64_t* bar = reinterpret_cast<uint64_t*>(
        (reinterpret_cast<uintptr_t>(sBitsB + Warp32Smem::bitwords(N)) + 7) & ~uintptr_t(7));
    // backward: v[i] = exp(-f[i]/sqrt(W)) for open cells, 0 otherwise (dense softmax numerator)
    float* sV = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(bar + 1) + 15) & ~uintptr_t(15));
    const int nv4 = ((N + 127) & ~127) >> 2;   // float4 count of the padded v plane (multiple of 32)

    const float* gCost = p.cost + int64_t(b) * p.cost_stride;
    const float* gStart = p.start + int64_t(b) * p.start_stride;
    const float* gGoal = p.goal + int64_t(b) * p.goal_stride;
    const float* gObst = p.obst + int64_t(b) * p.obst_stride;
    const bool obst_is_cost = (gObst == gCost);

    // ---------------- prologue: stage planes (TMA), build masks and h ----------------------
    const bool size_ok = ((N & 3) == 0);
    const bool t_cost = size_ok && aligned16(gCost);
    const bool t_start = size_ok && aligned16(gStart);
    const bool t_goal = size_ok && aligned16(gGoal);
    const bool t_obst = size_ok && aligned16(gObst) && !obst_is_cost;
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    uint32_t tx = 0;
    tx += stage_plane(sCost, gCost, N, t_cost, lane);
    tx += stage_plane(sF, gStart, N, t_start, lane);   // start plane parked in the f plane
    tx += stage_plane(sH, gGoal, N, t_goal, lane);     // goal plane parked in the h plane
    if (!obst_is_cost) tx += stage_plane(sG, gObst, N, t_obst, lane);  // obstacles parked in g
    if (tx) {
        if (lane == 0) {
            mbar_expect_tx(bar, tx);
            if (t_cost) tma_load_1d(sCost, gCost, uint32_t(N) * 4u, bar);
            if (t_start) tma_load_1d(sF, gStart, uint32_t(N) * 4u, bar);
            if (t_goal) tma_load_1d(sH, gGoal, uint32_t(N) * 4u, bar);
            if (t_obst) tma_load_1d(sG, gObst, uint32_t(N) * 4u, bar);
        }
        mbar_wait(bar, 0);
    }
    __syncwarp();

    const float* sObst = obst_is_cost ? sCost : sG;
    int start_idx = -1, goal_idx = -1;
    for (int w = 0; w < nwords; ++w) {
        const int i = (w << 5) + lane;
        const bool in = i < N;
        const float vo = in ? sObst[i] : 0.f;
        const float vs = in ? sF[i] : 0.f;
        const float vg = in ? sH[i] : 0.f;
        const uint32_t wo = __ballot_sync(kFull, vo != 0.f);
        const uint32_t ws = __ballot_sync(kFull, vs != 0.f);
        const uint32_t wg = __ballot_sync(kFull, vg != 0.f);
        if (lane == 0) sBitsA[w] = wo;
        if (start_idx < 0 && ws) start_idx = (w << 5) + __ffs(ws) - 1;
        if (goal_idx < 0 && wg) goal_idx = (w << 5) + __ffs(wg) - 1;
    }
    if (lane == 0) sBitsA[nwords] = 0u;
    if (goal_idx < 0) goal_idx = 0;  // argmax of an all-zero plane (differentiable_astar.py:197)
    __syncwarp();

    const uint32_t rowmask = (W == 32) ? 0xFFFFFFFFu : ((1u << W) - 1u);
    uint32_t pass = 0u;
    if (lane < H) {
        const int bit0 = lane * W;
        pass = __funnelshift_r(sBitsA[bit0 >> 5], sBitsA[(bit0 >> 5) + 1], bit0 & 31) & rowmask;
    }
    const int gy = goal_idx / W, gx = goal_idx - gy * W;
    __syncwarp();
    // h = heuristic + cost (differentiable_astar.py:191-192); overwrites the parked goal plane
    {
        int y = 0, x = lane;
        while (x >= W) { x -= W; ++y; }
        for (int i = lane; i < N; i += 32) {
            sH[i] = __fadd_rn(heuristic(y, x, gy, gx), sCost[i]);
            x += 32;
            while (x >= W) { x -= W; ++y; }
        }
    }
    __syncwarp();

    const float gr = p.g_ratio, omg = p.one_minus_g_ratio;
    uint32_t open = 0u, closed = 0u;
    uint32_t rm_key = kKeyInf;
    int rm_col = 0;
    if (lane == 0) sPar[goal_idx] = uint16_t(goal_idx);  // parents initialised to the goal index (:195-198)
    float Gh[32], acc[32];        // backward: upstream gradient / accumulator, cell 4*(lane+32j)+e
    int Tb = 0, ts_in = NASTAR_TS_CAPPED;
    if (kBwd) {
        float4* sV4 = reinterpret_cast<float4*>(sV);
        for (int q = lane; q < nv4; q += 32) sV4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        Tb = *a.T_batch;
        ts_in = a.t_solve_in[b];
        const float* gG = a.grad_hist + int64_t(b) * a.grad_stride;
        const bool vec = ((N & 3) == 0) && aligned16(gG);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i0 = 4 * (lane + 32 * j);
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vec) {
                if (i0 < N) g4 = __ldg(reinterpret_cast<const float4*>(gG + i0));
            } else {
                if (i0 + 0 < N) g4.x = __ldg(gG + i0 + 0);
                if (i0 + 1 < N) g4.y = __ldg(gG + i0 + 1);
                if (i0 + 2 < N) g4.z = __ldg(gG + i0 + 2);
                if (i0 + 3 < N) g4.w = __ldg(gG + i0 + 3);
            }
            // clamp(hist + sel) blocks the gradient at a goal that is re-selected after its solve
            // step (pre-clamp value 2, differentiable_astar.py:222-223)
            const bool blocked = (ts_in >= 0) && (ts_in < Tb - 1);
            Gh[4 * j + 0] = (blocked && i0 + 0 == goal_idx) ? 0.f : g4.x;
            Gh[4 * j + 1] = (blocked && i0 + 1 == goal_idx) ? 0.f : g4.y;
            Gh[4 * j + 2] = (blocked && i0 + 2 == goal_idx) ? 0.f : g4.z;
            Gh[4 * j + 3] = (blocked && i0 + 3 == goal_idx) ? 0.f : g4.w;
            acc[4 * j + 0] = acc[4 * j + 1] = acc[4 * j + 2] = acc[4 * j + 3] = 0.f;
        }
        __syncwarp();
    }
    if (start_idx >= 0) {
        const int sy = start_idx / W, sx = start_idx - sy * W;
        const float f0 = f_value(gr, omg, 0.f, sH[start_idx]);
        if (lane == 0) {
            sPar[start_idx] = uint16_t(goal_idx);
            sG[start_idx] = 0.f;                   // g = 0 (:193); only ever read for opened cells
            sF[start_idx] = f0;
            if (kBwd) sV[start_idx] = expf(__fdiv_rn(-f0, a.sqrt_w));   // :207
        }
        if (lane == sy) {
            open = 1u << sx;                       // open_maps = start_maps (:187)
            rm_key = fkey(f0);
            rm_col = sx;
        }
    }
    __syncwarp();

    // ---------------- the search loop (differentiable_astar.py:203-252) --------------------
    const int T = p.T;
    int t_solve = NASTAR_TS_CAPPED;
    int steps = 0;
    int32_t* trace = kTrace ? (p.trace + int64_t(b) * T) : nullptr;
    uint32_t* sOpenRow = sBitsB;   // shared copy of every lane's open row (rescan reads row r's)
    sOpenRow[lane] = open;
    __syncwarp();
    const int rowbase = lane * W;
    const int Tloop = kBwd ? Tb : T;
    // post-solve steps are stationary when g_ratio >= 0.5 (the goal keeps being re-selected and
    // nothing changes): the backward then adds them in one go
    const bool stationary_ok = (gr >= 0.5f);
    for (int t = 0; t < Tloop; ++t) {
        // -- select: lexicographic arg-min of (f key, row, col) with two REDUX.MINs -----------
        const uint32_t m = __reduce_min_sync(kFull, rm_key);
        if (m == kKeyInf) { t_solve = NASTAR_TS_EXHAUSTED; break; }
        if (kBwd) {
            // y_t = v / sum(v) over the open set at the START of step t; accumulate
            // y_t[p] * (Gh[p] - <Gh, y_t>)  (times the number of identical post-solve steps)
            const float4* sV4 = reinterpret_cast<const float4*>(sV);
            float4 v[8];
            float s_ = 0.f, d_ = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[j] = (lane + 32 * j < nv4) ? sV4[lane + 32 * j] : make_float4(0.f, 0.f, 0.f, 0.f);
                s_ += (v[j].x + v[j].y) + (v[j].z + v[j].w);
                d_ = fmaf(Gh[4 * j + 0], v[j].x, d_);
                d_ = fmaf(Gh[4 * j + 1], v[j].y, d_);
                d_ = fmaf(Gh[4 * j + 2], v[j].z, d_);
                d_ = fmaf(Gh[4 * j + 3], v[j].w, d_);
            }
            s_ = warp_sum(s_);
            d_ = warp_sum(d_);
            const bool last = stationary_ok && (ts_in >= 0) && (t == ts_in + 1);
            const float wgt = last ? float(Tb - t) : 1.f;
            const float inv = wgt / s_;
            const float dd = d_ / s_;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[4 * j + 0] = fmaf(v[j].x * inv, Gh[4 * j + 0] - dd, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(v[j].y * inv, Gh[4 * j + 1] - dd, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(v[j].z * inv, Gh[4 * j + 2] - dd, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(v[j].w * inv, Gh[4 * j + 3] - dd, acc[4 * j + 3]);
            }
            if (last) break;
        }
        const uint32_t selrc = __reduce_min_sync(kFull, (rm_key == m) ? uint32_t((lane << 5) | rm_col) : 0xFFFFFFFFu);
        const int r = int(selrc >> 5), c = int(selrc & 31u);
        const int ind = r * W + c;
        steps = t + 1;
        if (kTrace && lane == 0) trace[t] = ind;
        const bool solved = (ind == goal_idx);              // :219-220
        const uint32_t cbit = 1u << c;
        // -- rescan inputs for row r (pre-expansion open cells minus the selected one); stale f
        //    values of cells relaxed this step are upper bounds and the fresh keys are merged below
        const uint32_t open_r = (kBwd && solved) ? sOpenRow[r] : (sOpenRow[r] & ~cbit);
        const float frs = sF[ind - c + lane];
        const uint32_t rs_key = ((open_r >> lane) & 1u) ? fkey(frs) : kKeyInf;
        // -- closed/open update of the selected cell (:222-225) ------------------------------
        const int dr = lane - r;
        if (dr == 0) {
            closed |= cbit;
            if (!solved) open &= ~cbit;                     // the goal stays open once selected
            rm_key = kKeyInf;                               // this row's minimum is rebuilt below
            if (kBwd && !solved) sV[ind] = 0.f;             // left the open set: no softmax weight
        }
        // -- expansion: rows r-1..r+1, columns c-1..c+1 (:228-249), branch-free ---------------
        uint32_t win = (c == 0) ? 3u : (7u << (c - 1));
        if (dr == 0) win &= ~cbit;
        const bool near = (dr >= -1) & (dr <= 1);
        const uint32_t cand = near ? (win & pass) : 0u;
        float gn[3], hn[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int x = c - 1 + k;
            const bool on = (cand >> (x & 31)) & 1u;        // cand has no bit for x = -1 or 32
            gn[k] = on ? sG[rowbase + x] : 0.f;
            hn[k] = on ? sH[rowbase + x] : 0.f;
        }
        const float g2 = __fadd_rn(sG[ind], sCost[ind]);    // :234, cost of the SELECTED node
        const float ag = __fmul_rn(gr, g2);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int x = c - 1 + k;
            const uint32_t bit = 1u << (x & 31);
            const bool on = (cand & bit) != 0u;
            // :235-236  idx = (1-open)(1-hist) + open*(g > g2), masked by passable neighbours
            const bool upd = on & (((open & bit) != 0u) ? (gn[k] > g2) : ((closed & bit) == 0u));
            const float fn = __fadd_rn(ag, __fmul_rn(omg, hn[k]));
            const uint32_t key = fkey(fn);
            if (upd) {
                sG[rowbase + x] = g2;                        // :238
                sF[rowbase + x] = fn;
                sPar[rowbase + x] = uint16_t(ind);           // :246-249
                if (kBwd) sV[rowbase + x] = expf(__fdiv_rn(-fn, a.sqrt_w));   // :207
            }
            open |= upd ? bit : 0u;                          // :242
            const bool better = upd & ((key < rm_key) | ((key == rm_key) & (x < rm_col)));
            rm_key = better ? key : rm_key;
            rm_col = better ? x : rm_col;
        }
        if (!kBwd && solved) { t_solve = t; break; }        // :251-252 (per-map early exit, App. A.4)
        if (near) sOpenRow[lane] = open;
        // -- fold the rescan into lane r's cached minimum -------------------------------------
        const uint32_t mr = __reduce_min_sync(kFull, rs_key);
        const uint32_t mc = __reduce_min_sync(kFull, (rs_key == mr) ? uint32_t(lane) : 0xFFFFFFFFu);
        const bool take = (dr == 0) & ((mr < rm_key) | ((mr == rm_key) & (int(mc) < rm_col)));
        rm_key = take ? mr : rm_key;
        rm_col = take ? int(mc) : rm_col;
        __syncwarp();
    }
    __syncwarp();

    if (kBwd) {
        // dL/dcost = -(1-g_ratio)/sqrt(W) * acc   (h = heuristic + cost, f = g_ratio*g + (1-g_ratio)*h)
        const float coef = -omg / a.sqrt_w;
        float* gOut = a.grad_cost + int64_t(b) * N;
        const bool vec = ((N & 3) == 0) && aligned16(gOut);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i0 = 4 * (lane + 32 * j);
            if (vec) {
                if (i0 < N)
                    *reinterpret_cast<float4*>(gOut + i0) = make_float4(coef * acc[4 * j], coef * acc[4 * j + 1],
                                                                        coef * acc[4 * j + 2], coef * acc[4 * j + 3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (i0 + e < N) gOut[i0 + e] = coef * acc[4 * j + e];
            }
        }
        return;
    }

    // ---------------- backtrack (differentiable_astar.py:96-125, App. A.3) ------------------
    uint32_t path = 0u;
    {
        const int gyy = goal_idx / W;
        if (lane == gyy) path |= 1u << (goal_idx - gyy * W);
        int loc = sPar[goal_idx];
        const int hops = (t_solve >= 0) ? N : (T - 1);
        for (int k = 0; k < hops; ++k) {
            const int y = loc / W;
            if (lane == y) path |= 1u << (loc - y * W);
            if (loc == start_idx || loc == goal_idx) break;  // reached the start (or a self-loop)
            loc = sPar[loc];
        }
    }

    // ---------------- epilogue: coalesced stores of histories / paths -----------------------
    sBitsA[lane] = closed;
    sBitsB[lane] = path;
    __syncwarp();
    float* gHist = p.histories + int64_t(b) * N;
    long long* gPath = reinterpret_cast<long long*>(p.paths) + int64_t(b) * N;
    if ((W & 3) == 0 && aligned16(gHist) && aligned16(gPath)) {
        const int n4 = N >> 2;
        int y = 0, x = lane << 2;
        while (x >= W) { x -= W; ++y; }
        for (int i4 = lane; i4 < n4; i4 += 32) {
            const uint32_t cb = sBitsA[y] >> x, pb = sBitsB[y] >> x;
            float4 hv = make_float4((cb & 1u) ? 1.f : 0.f, (cb & 2u) ? 1.f : 0.f, (cb & 4u) ? 1.f : 0.f,
                                    (cb & 8u) ? 1.f : 0.f);
            reinterpret_cast<float4*>(gHist)[i4] = hv;
            longlong2 p0 = make_longlong2((pb & 1u) ? 1ll : 0ll, (pb & 2u) ? 1ll : 0ll);
            longlong2 p1 = make_longlong2((pb & 4u) ? 1ll : 0ll, (pb & 8u) ? 1ll : 0ll);
            reinterpret_cast<longlong2*>(gPath)[2 * i4] = p0;
            reinterpret_cast<longlong2*>(gPath)[2 * i4 + 1] = p1;
            x += 128;
            while (x >= W) { x -= W; ++y; }
        }
    } else {
        int y = 0, x = lane;
        while (x >= W) { x -= W; ++y; }
        for (int i = lane; i < N; i += 32) {
            gHist[i] = ((sBitsA[y] >> x) & 1u) ? 1.f : 0.f;
            gPath[i] = ((sBitsB[y] >> x) & 1u) ? 1ll : 0ll;
            x += 32;
            while (x >= W) { x -= W; ++y; }
        }
    }
    if (lane == 0) {
        if (p.t_solve) p.t_solve[b] = t_solve;
        if (p.n_steps) p.n_steps[b] = steps;
    }
}

}  // namespace nastar
