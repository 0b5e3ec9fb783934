// Latency microbenchmarks for the warp-collective chain used by the A* step (developer probe).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048

template <int MODE>
__global__ void chain(uint32_t* out, long long* cyc, uint32_t seed) {
    __shared__ uint32_t sm[2048];
    const int lane = threadIdx.x;
    for (int i = lane; i < 2048; i += 32) sm[i] = (i * 1103515245u + seed) & 2047u;
    __syncwarp();
    uint32_t x = seed + lane * 2654435761u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0) {            // redux.min chain
            x = __reduce_min_sync(0xFFFFFFFFu, x) + lane + i;
        } else if (MODE == 1) {     // ballot + ffs chain
            uint32_t b = __ballot_sync(0xFFFFFFFFu, (x & 1u) != 0);
            x = x * 3u + __ffs(b | 0x80000000u) + lane;
        } else if (MODE == 2) {     // shfl idx chain
            x = __shfl_sync(0xFFFFFFFFu, x, x & 31) + lane;
        } else if (MODE == 3) {     // LDS pointer chase
            x = sm[x & 2047u];
        } else if (MODE == 4) {     // 5-step butterfly min
            uint32_t v = x;
            for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
            x = v + lane + i;
        } else if (MODE == 5) {     // redux + ballot + ffs + shfl (current selection)
            uint32_t m = __reduce_min_sync(0xFFFFFFFFu, x);
            int r = __ffs(__ballot_sync(0xFFFFFFFFu, x == m)) - 1;
            uint32_t c = __shfl_sync(0xFFFFFFFFu, x >> 3, r);
            x = x * 5u + c + lane + i;
        } else if (MODE == 6) {     // redux + redux (two-stage key/index)
            uint32_t m = __reduce_min_sync(0xFFFFFFFFu, x);
            uint32_t idx = __reduce_min_sync(0xFFFFFFFFu, x == m ? (lane << 5 | (x & 31)) : 0xFFFFu);
            x = x * 5u + idx + lane + i;
        } else if (MODE == 7) {     // FADD dependent chain (ALU baseline)
            float f = __uint_as_float(x);
            f = __fadd_rn(f, 1.0f);
            x = __float_as_uint(f);
        } else if (MODE == 8) {     // STS then LDS same address by another lane (smem round trip)
            sm[lane] = x;
            __syncwarp();
            x = sm[(lane + 1) & 31] + i;
            __syncwarp();
        } else if (MODE == 9) {     // match_any
            x = __match_any_sync(0xFFFFFFFFu, x & 3u) + x * 7u + i;
        } else if (MODE == 10) {    // redux.min on signed + uniform broadcast via shfl(0)
            x = __shfl_sync(0xFFFFFFFFu, x, 0) + lane + i;
        } else if (MODE == 11) {    // vote only
            uint32_t b = __ballot_sync(0xFFFFFFFFu, (x & 1u) != 0);
            x = x * 3u + b + lane;
        } else if (MODE == 12) {    // ffs only (brev+flo)
            x = x * 3u + __ffs(x | 0x80000000u) + lane;
        }
    }
    long long t1 = clock64();
    out[lane] = x;
    if (lane == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    chain<MODE><<<1, 32>>>(out, cyc, 12345u);
    chain<MODE><<<1, 32>>>(out, cyc, 999u);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %8.1f cycles/iter\n", name, double(h) / ITERS);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<7>("FADD chain (ALU baseline)");
    run<0>("redux.min.u32 (+iadd)");
    run<11>("ballot (+imad)");
    run<12>("ffs (+imad)");
    run<1>("ballot+ffs");
    run<2>("shfl.idx");
    run<10>("shfl.idx lane0 broadcast");
    run<3>("LDS pointer chase");
    run<4>("5-step shfl_xor butterfly min");
    run<5>("redux+ballot+ffs+shfl");
    run<6>("redux+redux");
    run<8>("STS->syncwarp->LDS->syncwarp");
    run<9>("match_any");
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
