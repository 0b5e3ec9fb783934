/*
 * nastar_b200.h — C ABI of the B200-native differentiable-A* engine (libnastar_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has no FFI: the
 * seam is the Python attribute `self.astar = DifferentiableAstar(...)`
 * (/root/reference/src/neural_astar/planner/astar.py:41-44,141-144) whose `.forward`
 * (/root/reference/src/neural_astar/planner/differentiable_astar.py:150-267) is replaced.
 * Each entry point below names the reference code it replaces.  All pointers are DEVICE
 * pointers unless stated otherwise; no torch types cross this boundary.
 *
 * Plane layout: every "plane" argument is B maps of H*W contiguous fp32 cells (row-major,
 * flat index = y*W + x, the same flat index the reference uses for `parents`,
 * differentiable_astar.py:195-198).  `*_stride` is the distance between consecutive maps
 * in ELEMENTS (so a [B,C,H,W] tensor's channel 0 is passed with stride C*H*W, matching
 * `cost_maps[:, 0]`, differentiable_astar.py:177-180).
 */
#ifndef NASTAR_B200_H_
#define NASTAR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NASTAR_B200_ABI_VERSION 2

/* status codes (the reference only has Python asserts, differentiable_astar.py:147,172-175) */
enum {
    NASTAR_OK = 0,
    NASTAR_EINVAL = 1,      /* bad shape / null pointer / T < 1 */
    NASTAR_EUNSUPPORTED = 2,/* map shape outside what this build handles */
    NASTAR_ECUDA = 3,       /* a CUDA runtime call failed; see nastar_b200_last_cuda_error() */
    NASTAR_EWORKSPACE = 4   /* workspace too small */
};

/* per-map outcome written to t_solve[] */
#define NASTAR_TS_CAPPED (-1)    /* step cap T reached before the goal was selected */
#define NASTAR_TS_EXHAUSTED (-2) /* open list ran empty: goal unreachable (the reference
                                    produces NaN -> IndexError here) */

/* nastar_fwd_params.flags */
#define NASTAR_FWD_NO_EARLY_EXIT 1 /* run exactly T steps per map: a solved map keeps stepping (re-selecting its
                                      goal, or other nodes when g_ratio < 0.5) like the reference's batch-synchronous
                                      loop does until the slowest map is solved (differentiable_astar.py:251-252).
                                      Per-map early exit is exact for g_ratio >= 0.5; the host
                                      side uses this flag to reproduce the batch coupling for g_ratio < 0.5 */
#define NASTAR_FWD_PAIR 2          /* validation pair in ONE launch (utils/training.py:63-87 runs the learned-cost
                                      search and VanillaAstar on the same batch): maps 0..B-1 are searched on
                                      `cost` (any cost_kind), maps B..2B-1 are the same problems searched with
                                      cost = obst (astar.py:93-94).  Every output array then holds 2B entries.
                                      Engine-1 shapes (H,W <= 32) run both halves in one kernel launch; other
                                      shapes issue the two halves back to back inside this call */

/* nastar_fwd_params.cost_kind: what `cost` points to (encoder -> search hand-off) */
#define NASTAR_COST_PLANE 0  /* finished cost maps, fp32 [B][H*W] (differentiable_astar.py:150-157) */
#define NASTAR_COST_LOGIT 1  /* the encoder's raw 1-channel output x, fp32 [B][H*W]; the kernel prologue applies
                                encoder.py:32-34: cost = sigmoid(x) * cost_scale */
#define NASTAR_COST_TAPS  2  /* per-pixel partial products of the encoder's last 3x3 conv (one output channel),
                                fp32 [B][H*W][9] (tap k = ky*3+kx innermost): the prologue gathers
                                x[y][x] = cost_bias + sum_k taps[y+ky-1][x+kx-1][k] (zero padding), then
                                cost = sigmoid(x) * cost_scale.  `cost_stride` is in elements of this layout
                                (9*H*W for a dense batch).  LOGIT / TAPS: engines 1 and 4 (H,W <= 64),
                                NASTAR_EUNSUPPORTED for larger shapes */

typedef struct nastar_fwd_params {
    /* inputs — differentiable_astar.py:150-157 (cost_maps, start_maps, goal_maps, obstacles_maps) */
    const float *cost;   int64_t cost_stride;
    const float *start;  int64_t start_stride;
    const float *goal;   int64_t goal_stride;
    const float *obst;   int64_t obst_stride;   /* may alias cost (VanillaAstar, astar.py:93-94) */
    int32_t B, H, W;
    /* scalars prepared on the host exactly as Python does (differentiable_astar.py:206):
       g_ratio and (1 - g_ratio) evaluated in double, then rounded to fp32 */
    float g_ratio;
    float one_minus_g_ratio;
    /* loop bound int(Tmax_eff * W * W), differentiable_astar.py:200-202 */
    int32_t T;
    /* NASTAR_FWD_* flags */
    int32_t flags;
    /* NASTAR_COST_*; cost_scale = the encoder's `const` (encoder.py:25-28,34), cost_bias = folded bias of the
       last conv (NASTAR_COST_TAPS only).  Zero-initialised params mean a plain cost plane */
    int32_t cost_kind;
    float   cost_scale;
    float   cost_bias;
    /* outputs */
    float   *histories;  /* [B][H*W] fp32 in {0,1}   — AstarOutput.histories, :265 */
    int64_t *paths;      /* [B][H*W] int64 in {0,1}  — AstarOutput.paths (backtrack, :96-125) */
    int32_t *t_solve;    /* [B] step at which the goal was selected, or NASTAR_TS_*; nullable */
    int32_t *n_steps;    /* [B] number of selection steps executed for the map; nullable */
    int32_t *trace;      /* [B][T] selected flat index per step (-1 beyond n_steps); nullable.
                            Feeds store_intermediate_results (:210-216) */
    int32_t *n_closed;   /* [B] number of closed cells = histories.sum() per map; nullable.  With path_len it
                            replaces the host-side sums of the validation metrics (utils/training.py:71-85) */
    int32_t *path_len;   /* [B] number of cells on the path = paths.sum() per map; nullable */
    /* scratch for maps whose state does not fit in shared memory; see
       nastar_b200_forward_workspace_bytes().  nullable when that returns 0 */
    void    *workspace;
    size_t   workspace_bytes;
} nastar_fwd_params;

typedef struct nastar_bwd_params {
    /* the forward's inputs again (the backward replays the deterministic search) */
    const float *cost;   int64_t cost_stride;
    const float *start;  int64_t start_stride;
    const float *goal;   int64_t goal_stride;
    const float *obst;   int64_t obst_stride;
    int32_t B, H, W;
    float g_ratio;
    float one_minus_g_ratio;
    float sqrt_w;             /* fl32(sqrt(W)), differentiable_astar.py:207 */
    /* number of loop iterations the reference would have executed for this batch:
       T_batch = min(T, 1 + max_b t_solve[b]) (batch-coupled stop, :251-252).  Read on the
       DEVICE from *T_batch so that no host synchronisation is needed between forward
       and backward */
    const int32_t *T_batch;
    const int32_t *t_solve;   /* [B] the forward's t_solve[] for the same inputs and loop bound.  Used for (a) the
                                 stationary post-solve shortcut when g_ratio >= 0.5 and (b) the goal clamp: the
                                 gradient at a map's goal is blocked iff 0 <= t_solve[b] < T_batch-1 (the goal is
                                 then selected again and clamp(hist+sel) saw 2, differentiable_astar.py:222-223).
                                 For g_ratio < 0.5 a solved map need not re-select its goal: pass 0 for maps
                                 whose goal is selected at least twice within T_batch steps and T_batch otherwise
                                 (the Python module derives this from a NO_EARLY_EXIT trace) */
    const float *grad_histories; int64_t grad_stride; /* dL/d histories, [B][H*W] */
    float *grad_cost;                                  /* dL/d cost_maps, [B][H*W], overwritten */
    void  *workspace;
    size_t workspace_bytes;
} nastar_bwd_params;

/* ABI version of the loaded library (== NASTAR_B200_ABI_VERSION at build time). */
int nastar_b200_abi_version(void);

/* Bytes of device scratch nastar_b200_forward needs for this shape: 0 for H,W <= 64 (every 32x32 / 12x12 / 64x64
 * config); larger maps need the work queue + per-map flags of the binary-cost engine and, beyond 128x128, the
 * per-CTA state slots of the generic engine.  Passing less than this but at least the generic engine's own need
 * still works (the binary-cost fast path is then skipped). */
size_t nastar_b200_forward_workspace_bytes(int32_t B, int32_t H, int32_t W);
size_t nastar_b200_backward_workspace_bytes(int32_t B, int32_t H, int32_t W);

/* Replaces DifferentiableAstar.forward's loop + backtrack (differentiable_astar.py:187-255).
 * Asynchronous on `stream` (a cudaStream_t passed as void*). Returns NASTAR_*. */
int nastar_b200_forward(const nastar_fwd_params *p, void *stream);

/* Replaces autograd through the loop (closed form): writes dL/dcost. */
int nastar_b200_backward(const nastar_bwd_params *p, void *stream);

/* T_batch[0] = min(T, 1 + max_b n_steps-style stop) computed on the device from the forward's
 * t_solve/n_steps arrays (replaces the host-synchronising torch.all(...) of :251). */
int nastar_b200_batch_steps(const int32_t *t_solve, const int32_t *n_steps, int32_t B, int32_t T,
                            int32_t *T_batch, void *stream);

/* Which forward engine a shape dispatches to: 1 = warp-resident (H,W <= 32), 4 = warp-resident 64-wide
 * (H,W <= 64), 2 = generic warp engine with shared-memory state, 3 = generic with global-memory state,
 * 0 = unsupported.  Engine-2/3 shapes whose cost
 * plane IS the obstacle plane (same pointer and stride: VanillaAstar, astar.py:93-94) are first offered to
 * engine 5, the CTA-per-map binary-cost engine (csrc/nastar_bin16.cuh); maps it cannot take (a cost value
 * outside {0,1}) are re-run by engine 2/3 inside the same call. */
int nastar_b200_engine_for(int32_t H, int32_t W);

/* 1 if engine 5 (binary-cost, whole map in one CTA's shared memory) can hold this shape. */
int nastar_b200_bin16_supported(int32_t H, int32_t W);

/* Encoder input assembly, replaces NeuralAstar.encode's glue (astar.py:172-177): out is the channels-last
 * (NHWC) tensor [B][Hm][Wm][C+1] with channels 0..C-1 = map_designs[b][c] (fp32 NCHW [B][C][Hm][Wm], contiguous)
 * and channel C = (start + goal) nearest-upsampled from [H][W] to [Hm][Wm] (F.interpolate mode="nearest";
 * identity when the sizes agree).  start/goal are [B][H*W] planes with element strides. */
int nastar_b200_pack_inputs(const float *map_designs, int32_t C, int32_t Hm, int32_t Wm,
                            const float *start, int64_t start_stride, const float *goal, int64_t goal_stride,
                            int32_t B, int32_t H, int32_t W, float *out, void *stream);

/* cost = sigmoid(bias + 9-tap gather of `taps`) * scale for every cell, written as a dense [B][H*W] plane — the
 * same arithmetic, in the same order, as nastar_b200_forward's NASTAR_COST_TAPS prologue (so `encode()` and the
 * fused forward agree bit for bit).  taps: fp32 [B][H*W][9]. */
int nastar_b200_cost_from_taps(const float *taps, int32_t B, int32_t H, int32_t W, float bias, float scale,
                               float *cost, void *stream);

/* First encoder layer of the "m+" planners fused with the input assembly (astar.py:172-177 + the first
 * Conv2d(2,32,3,padding=1) + BatchNorm + ReLU of encoder.py:60-78): out[b][y][x][0..31] (channels-last fp32) =
 * relu(bias + conv3x3(cat(map_designs, start + goal))).  map_designs: fp32 [B][H*W] contiguous (one channel);
 * start / goal: [B][H*W] planes with element strides; w_host: HOST pointer, fp32 [9 taps][2 in][32 out] with
 * BatchNorm folded in; bias_host: HOST pointer, fp32 [32] (both copied into the launch's parameter block, i.e. the
 * constant bank — safe to capture in a CUDA graph).  Same spatial size for maps and marks only; B*H*W <= 2^30. */
int nastar_b200_conv1_marks(const float *map_designs, const float *start, int64_t start_stride, const float *goal,
                            int64_t goal_stride, int32_t B, int32_t H, int32_t W, const float *w_host,
                            const float *bias_host, float *out, void *stream);

/* Head of the encoder (its last 3x3 conv has ONE output channel, encoder.py:60-78): taps[p][k] = sum_c x[p][c] *
 * w[c][k] for every pixel p of the channels-last activation x (fp32 [P][C], C in {32,64,128,256}), k = ky*3+kx.
 * `w` is a HOST pointer to C*9 floats ([c][k], BatchNorm folded): it is passed to the kernel by value (constant
 * bank), so it may change between calls without any device copy.  The 9-tap spatial gather + bias + sigmoid * const
 * that finish the layer happen in nastar_b200_forward's NASTAR_COST_TAPS prologue (or nastar_b200_cost_from_taps). */
int nastar_b200_head_taps(const float *x, int64_t P, int32_t C, const float *w_host, float *taps, void *stream);

/* Self-test of engine 5's branch-free square root: counts, on the device, the integers i in [0, n) for which it
 * differs in any bit from the IEEE-rounded sqrtf(i) that get_heuristic needs (differentiable_astar.py:47-50).
 * `mismatches` is a device int the caller zeroed.  n = 2*511*511+1 covers every argument a 512-row map can produce. */
int nastar_b200_selftest_sqrt(int32_t n, int32_t *mismatches, void *stream);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t nastar_b200_launch_count(void);

const char *nastar_b200_status_string(int status);
/* cudaGetErrorString of the last failing CUDA call made by this library (thread-unsafe, debug aid) */
const char *nastar_b200_last_cuda_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NASTAR_B200_H_ */
