// nastar_api.cu — C-ABI entry points of libnastar_b200.so (see include/nastar_b200.h).
// Host side only validates, picks an engine and launches; no torch types, no CPU fallback.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>

#include "../../include/nastar_b200.h"
#include "nastar_generic.cuh"
#include "nastar_warp32.cuh"
#include "nastar_warp64.cuh"

namespace {
std::atomic<uint64_t> g_launches{0};
cudaError_t g_last_err = cudaSuccess;

constexpr size_t kMaxDynSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100
constexpr int kGlobalCtasPerSm = 8;

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// one-time fill of the 32x32 heuristic table on the current device (stream-ordered before first use)
cudaError_t ensure_heur32(cudaStream_t stream) {
    static std::atomic<uint64_t> done_mask{0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = uint64_t(1) << (dev & 63);
    if (done_mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    nastar::heur32_init_kernel<<<4, 256, 0, stream>>>();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // later launches on OTHER streams must also see the table: finish the fill before publishing
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    done_mask.fetch_or(bit, std::memory_order_release);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaSuccess;
}

cudaError_t ensure_heur64(cudaStream_t stream) {
    static std::atomic<uint64_t> done_mask{0};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t bit = uint64_t(1) << (dev & 63);
    if (done_mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    nastar::heur64_init_kernel<<<16, 256, 0, stream>>>();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    done_mask.fetch_or(bit, std::memory_order_release);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return cudaSuccess;
}

// generic engine variant for a shape: 2 = state in shared memory, 3 = state in the HBM workspace, 0 = too large
int generic_engine_for(int32_t H, int32_t W) {
    if (H <= 0 || W <= 0 || int64_t(H) * W > (int64_t(1) << 30)) return 0;
    const nastar::GenericLayout L(H, W);
    if (L.smem_common() + L.smem_planes() <= kMaxDynSmem) return 2;
    if (L.smem_common() <= kMaxDynSmem) return 3;
    return 0;
}

int generic_slots(int B) {
    const int cap = num_sms() * kGlobalCtasPerSm;
    return B < cap ? B : cap;
}

inline int cuda_fail(cudaError_t e) {
    g_last_err = e;
    return NASTAR_ECUDA;
}

__global__ void batch_steps_kernel(const int32_t* __restrict__ t_solve, const int32_t* __restrict__ n_steps,
                                   int B, int T, int32_t* __restrict__ out) {
    // T_batch = number of iterations of the reference's batch-synchronous loop
    // (differentiable_astar.py:203,251-252): it stops right after the slowest map's solve step,
    // or runs all T iterations if some map never reaches its goal.
    int m = 0;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const int ts = t_solve[i];
        const int v = (ts >= 0) ? (ts + 1) : T;
        m = max(m, v);
        (void)n_steps;
    }
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    __shared__ int sm[32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0;
        for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        if (threadIdx.x == 0) out[0] = min(m, T);
    }
}
}  // namespace

extern "C" {

int nastar_b200_abi_version(void) { return NASTAR_B200_ABI_VERSION; }

int nastar_b200_engine_for(int32_t H, int32_t W) {
    if (H <= 0 || W <= 0) return 0;
    if (H <= 32 && W <= 32) return 1;
    if (H <= 64 && W <= 64) return 4;   // forward; the backward of these shapes runs on the generic engine
    return generic_engine_for(H, W);
}

size_t nastar_b200_forward_workspace_bytes(int32_t B, int32_t H, int32_t W) {
    if (B <= 0 || nastar_b200_engine_for(H, W) != 3) return 0;
    return size_t(generic_slots(B)) * nastar::GenericLayout(H, W).slot_bytes();
}

size_t nastar_b200_backward_workspace_bytes(int32_t B, int32_t H, int32_t W) {
    if (B <= 0 || nastar_b200_engine_for(H, W) == 1) return 0;
    const int e = generic_engine_for(H, W);
    if (e == 0) return 0;
    return size_t(generic_slots(B)) * nastar::GenericLayout(H, W).slot_total(e == 3, true);
}

int nastar_b200_forward(const nastar_fwd_params* p, void* stream_v) {
    if (!p || !p->cost || !p->start || !p->goal || !p->obst || !p->histories || !p->paths) return NASTAR_EINVAL;
    if (p->B <= 0 || p->H <= 0 || p->W <= 0 || p->T < 1) return NASTAR_EINVAL;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const int engine = nastar_b200_engine_for(p->H, p->W);
    if (engine == 0) return NASTAR_EUNSUPPORTED;
    if (p->trace) {
        cudaError_t e = cudaMemsetAsync(p->trace, 0xFF, size_t(p->B) * size_t(p->T) * sizeof(int32_t), stream);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    if (engine == 1) {
        cudaError_t he = ensure_heur32(stream);
        if (he != cudaSuccess) return cuda_fail(he);
        nastar::W32Args a{};
        a.f = *p;
        const bool noexit = (p->flags & NASTAR_FWD_NO_EARLY_EXIT) != 0;
        if (p->trace) {
            if (noexit) nastar::astar_warp32_kernel<true, false, true><<<p->B, 32, 0, stream>>>(a);
            else nastar::astar_warp32_kernel<true, false, false><<<p->B, 32, 0, stream>>>(a);
        } else {
            if (noexit) nastar::astar_warp32_kernel<false, false, true><<<p->B, 32, 0, stream>>>(a);
            else nastar::astar_warp32_kernel<false, false, false><<<p->B, 32, 0, stream>>>(a);
        }
        g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if (engine == 4) {
        cudaError_t he = ensure_heur64(stream);
        if (he != cudaSuccess) return cuda_fail(he);
        const bool noexit = (p->flags & NASTAR_FWD_NO_EARLY_EXIT) != 0;
        const size_t smem = sizeof(nastar::W64Smem);
        auto launch = [&](auto kernel) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            kernel<<<p->B, 32, smem, stream>>>(*p);
            return cudaSuccess;
        };
        cudaError_t e;
        if (p->trace) e = noexit ? launch(nastar::astar_warp64_kernel<true, true>) : launch(nastar::astar_warp64_kernel<true, false>);
        else e = noexit ? launch(nastar::astar_warp64_kernel<false, true>) : launch(nastar::astar_warp64_kernel<false, false>);
        if (e != cudaSuccess) return cuda_fail(e);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    } else {
        const nastar::GenericLayout L(p->H, p->W);
        const bool global = (engine == 3);
        const size_t smem = L.smem_common() + (global ? 0 : L.smem_planes());
        int grid = p->B;
        if (global) {
            grid = generic_slots(p->B);
            if (!p->workspace || p->workspace_bytes < size_t(grid) * L.slot_bytes()) return NASTAR_EWORKSPACE;
        }
        nastar::GenArgs ga{};
        ga.f = *p;
        auto launch = [&](auto kernel) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            kernel<<<grid, 32, smem, stream>>>(ga);
            return cudaSuccess;
        };
        const bool noexit = (p->flags & NASTAR_FWD_NO_EARLY_EXIT) != 0;
        cudaError_t e;
        if (noexit) {
            if (global) e = p->trace ? launch(nastar::astar_generic_kernel<true, true, false, true>) : launch(nastar::astar_generic_kernel<true, false, false, true>);
            else e = p->trace ? launch(nastar::astar_generic_kernel<false, true, false, true>) : launch(nastar::astar_generic_kernel<false, false, false, true>);
        } else {
            if (global) e = p->trace ? launch(nastar::astar_generic_kernel<true, true, false>) : launch(nastar::astar_generic_kernel<true, false, false>);
            else e = p->trace ? launch(nastar::astar_generic_kernel<false, true, false>) : launch(nastar::astar_generic_kernel<false, false, false>);
        }
        if (e != cudaSuccess) return cuda_fail(e);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    return NASTAR_OK;
}

int nastar_b200_backward(const nastar_bwd_params* p, void* stream_v) {
    if (!p || !p->cost || !p->start || !p->goal || !p->obst || !p->grad_histories || !p->grad_cost || !p->T_batch ||
        !p->t_solve)
        return NASTAR_EINVAL;
    if (p->B <= 0 || p->H <= 0 || p->W <= 0) return NASTAR_EINVAL;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    int engine = nastar_b200_engine_for(p->H, p->W);
    if (engine != 1) engine = generic_engine_for(p->H, p->W);   // warp64 shapes: backward on the generic engine
    if (engine == 0) return NASTAR_EUNSUPPORTED;
    if (engine >= 2) {
        const nastar::GenericLayout L(p->H, p->W);
        const bool global = (engine == 3);
        const int grid = generic_slots(p->B);
        if (!p->workspace || p->workspace_bytes < size_t(grid) * L.slot_total(global, true)) return NASTAR_EWORKSPACE;
        nastar::GenArgs ga{};
        ga.f.cost = p->cost;   ga.f.cost_stride = p->cost_stride;
        ga.f.start = p->start; ga.f.start_stride = p->start_stride;
        ga.f.goal = p->goal;   ga.f.goal_stride = p->goal_stride;
        ga.f.obst = p->obst;   ga.f.obst_stride = p->obst_stride;
        ga.f.B = p->B; ga.f.H = p->H; ga.f.W = p->W;
        ga.f.g_ratio = p->g_ratio;
        ga.f.one_minus_g_ratio = p->one_minus_g_ratio;
        ga.f.workspace = p->workspace;
        ga.f.workspace_bytes = p->workspace_bytes;
        ga.sqrt_w = p->sqrt_w;
        ga.T_batch = p->T_batch;
        ga.t_solve_in = p->t_solve;
        ga.grad_hist = p->grad_histories;
        ga.grad_stride = p->grad_stride;
        ga.grad_cost = p->grad_cost;
        const size_t smem = L.smem_common() + (global ? 0 : L.smem_planes());
        cudaError_t e;
        if (global) {
            e = cudaFuncSetAttribute(nastar::astar_generic_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e == cudaSuccess) nastar::astar_generic_kernel<true, false, true><<<grid, 32, smem, stream>>>(ga);
        } else {
            e = cudaFuncSetAttribute(nastar::astar_generic_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e == cudaSuccess) nastar::astar_generic_kernel<false, false, true><<<grid, 32, smem, stream>>>(ga);
        }
        if (e != cudaSuccess) return cuda_fail(e);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e);
        return NASTAR_OK;
    }
    {
        cudaError_t he = ensure_heur32(stream);
        if (he != cudaSuccess) return cuda_fail(he);
    }
    nastar::W32Args a{};
    a.f.cost = p->cost;   a.f.cost_stride = p->cost_stride;
    a.f.start = p->start; a.f.start_stride = p->start_stride;
    a.f.goal = p->goal;   a.f.goal_stride = p->goal_stride;
    a.f.obst = p->obst;   a.f.obst_stride = p->obst_stride;
    a.f.B = p->B; a.f.H = p->H; a.f.W = p->W;
    a.f.g_ratio = p->g_ratio;
    a.f.one_minus_g_ratio = p->one_minus_g_ratio;
    a.f.T = 0;  // the loop bound comes from *T_batch on the device
    a.sqrt_w = p->sqrt_w;
    a.T_batch = p->T_batch;
    a.t_solve_in = p->t_solve;
    a.grad_hist = p->grad_histories;
    a.grad_stride = p->grad_stride;
    a.grad_cost = p->grad_cost;
    // dynamic shared memory = the dense softmax-numerator plane v (padded 32x32 fp32)
    nastar::astar_warp32_kernel<false, true><<<p->B, 32, nastar::kCells * sizeof(float), stream>>>(a);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    return NASTAR_OK;
}

int nastar_b200_batch_steps(const int32_t* t_solve, const int32_t* n_steps, int32_t B, int32_t T, int32_t* T_batch,
                            void* stream_v) {
    if (!t_solve || !T_batch || B <= 0 || T < 1) return NASTAR_EINVAL;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    batch_steps_kernel<<<1, 256, 0, stream>>>(t_solve, n_steps, B, T, T_batch);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    return NASTAR_OK;
}

uint64_t nastar_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* nastar_b200_status_string(int s) {
    switch (s) {
        case NASTAR_OK: return "ok";
        case NASTAR_EINVAL: return "invalid argument";
        case NASTAR_EUNSUPPORTED: return "unsupported map shape or feature";
        case NASTAR_ECUDA: return "CUDA runtime error";
        case NASTAR_EWORKSPACE: return "workspace too small";
        default: return "unknown status";
    }
}

const char* nastar_b200_last_cuda_error(void) { return cudaGetErrorString(g_last_err); }

}  // extern "C"
