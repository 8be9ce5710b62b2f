// ubench.cu -- warp-primitive throughput on B200 (development aid, run under gpurun).
// Prints SM cycles per warp-level operation at 32 resident warps per SM, so that the ranking
// step of the radix kernels can be budgeted against the HBM-bound cycle count per warp item.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef uint32_t u32;
typedef uint64_t u64;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ u32 lcg(u32 &s) { s = s * 1664525u + 1013904223u; return s >> 12; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

template <int BITS>
__device__ __forceinline__ unsigned ballot_peers(u32 d) {
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
        bool bit = (d >> b) & 1u;
        unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

enum Mode { BASE = 0, MATCH, BALLOT8, BALLOT10, SHFL, ATOMS_ALL, ATOMS_LEADER, LDSSTS, RANK_MATCH, RANK_BALLOT8, RANK_BALLOT10, RANK_BALLOT8_X4, RANK_MATCH_X4, NMODES };
static const char *mode_name[] = {"base(lcg only)", "match.any", "ballot x8 peers", "ballot x10 peers", "shfl.idx", "atoms all lanes (256 bins)",
                                  "match + atoms leaders", "lds+sts random", "rank: match+atom+shfl", "rank: ballot8+lds/sts+shfl", "rank: ballot10+lds/sts+shfl (1024 bins)",
                                  "rank: ballot8 x4 interleaved", "rank: match x4 interleaved"};

template <int MODE>
__global__ void __launch_bounds__(256) k(u32 *out, int iters) {
    __shared__ u32 hist[8][1024];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < 1024; i += 32) hist[warp][i] = 0;
    __syncwarp();
    u32 s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1;
    u32 acc = 0;
    const unsigned lt = lanemask_lt();
    u32 *h = hist[warp];
    if (MODE == RANK_BALLOT8_X4 || MODE == RANK_MATCH_X4) {
        for (int it = 0; it < iters; it += 4) {
            u32 d[4]; unsigned peers[4]; u32 before[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) d[q] = lcg(s) & 255u;
#pragma unroll
            for (int q = 0; q < 4; ++q) peers[q] = MODE == RANK_MATCH_X4 ? __match_any_sync(0xffffffffu, d[q]) : ballot_peers<8>(d[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                before[q] = 0;
                if ((peers[q] & lt) == 0) { before[q] = h[d[q]]; h[d[q]] = before[q] + __popc(peers[q]); }
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += __shfl_sync(0xffffffffu, before[q], __ffs(peers[q]) - 1) + __popc(peers[q] & lt);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
            u32 r = lcg(s);
            u32 d = r & 255u;
            if (MODE == BASE) acc += d;
            if (MODE == MATCH) acc += __match_any_sync(0xffffffffu, d);
            if (MODE == BALLOT8) acc += ballot_peers<8>(d);
            if (MODE == BALLOT10) acc += ballot_peers<10>(r & 1023u);
            if (MODE == SHFL) acc += __shfl_sync(0xffffffffu, r, d & 31);
            if (MODE == ATOMS_ALL) acc += atomicAdd(&h[d], 1u);
            if (MODE == ATOMS_LEADER) {
                unsigned p = __match_any_sync(0xffffffffu, d);
                if ((p & lt) == 0) acc += atomicAdd(&h[d], (u32)__popc(p));
                __syncwarp();
            }
            if (MODE == LDSSTS) { u32 v = h[d]; __syncwarp(); h[(d + 1) & 255] = v + 1; __syncwarp(); acc += v; }
            if (MODE == RANK_MATCH) {
                unsigned p = __match_any_sync(0xffffffffu, d);
                u32 before = 0;
                if ((p & lt) == 0) before = atomicAdd(&h[d], (u32)__popc(p));
                __syncwarp();
                acc += __shfl_sync(0xffffffffu, before, __ffs(p) - 1) + __popc(p & lt);
            }
            if (MODE == RANK_BALLOT8 || MODE == RANK_BALLOT10) {
                if (MODE == RANK_BALLOT10) d = r & 1023u;
                unsigned p = MODE == RANK_BALLOT8 ? ballot_peers<8>(d) : ballot_peers<10>(d);
                u32 before = 0;
                if ((p & lt) == 0) { before = h[d]; h[d] = before + __popc(p); }
                __syncwarp();
                acc += __shfl_sync(0xffffffffu, before, __ffs(p) - 1) + __popc(p & lt);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// is the order in which one ATOMS instruction serves lanes that hit the same address the lane order?
__global__ void atoms_order_kernel(u32 *viol, int iters) {
    __shared__ u32 hist[8][256];
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 7;
    const unsigned lt = lanemask_lt();
    u32 bad = 0;
    for (int it = 0; it < iters; ++it) {
        for (int i = lane; i < 256; i += 32) hist[warp][i] = 0;
        __syncwarp();
        u32 d = lcg(s) & ((it & 1) ? 255u : 7u);
        unsigned p = __match_any_sync(0xffffffffu, d);
        u32 got = atomicAdd(&hist[warp][d], 1u);
        if (got != (u32)__popc(p & lt)) ++bad;
        __syncwarp();
    }
    if (bad) atomicAdd(viol, bad);
}

template <int MODE>
static void run(u32 *out, int iters) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int ctas = 148 * 4;
    k<MODE><<<ctas, 256>>>(out, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k<MODE><<<ctas, 256>>>(out, iters);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    // per SM: 32 warps each doing `iters` ops
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-45s %8.3f ms  %7.2f SM-cycles per warp-op (32 warps/SM)\n", mode_name[MODE], ms, cyc / (32.0 * iters));
}

int main() {
    u32 *out; CK(cudaMalloc(&out, 148 * 4 * 256 * 4));
    const int iters = 20000;
    run<BASE>(out, iters); run<MATCH>(out, iters); run<BALLOT8>(out, iters); run<BALLOT10>(out, iters); run<SHFL>(out, iters);
    run<ATOMS_ALL>(out, iters); run<ATOMS_LEADER>(out, iters); run<LDSSTS>(out, iters); run<RANK_MATCH>(out, iters);
    run<RANK_BALLOT8>(out, iters); run<RANK_BALLOT10>(out, iters); run<RANK_BALLOT8_X4>(out, iters); run<RANK_MATCH_X4>(out, iters);
    u32 *viol; CK(cudaMalloc(&viol, 4)); CK(cudaMemset(viol, 0, 4));
    atoms_order_kernel<<<148 * 4, 256>>>(viol, 20000);
    u32 hv; CK(cudaMemcpy(&hv, viol, 4, cudaMemcpyDeviceToHost));
    printf("ATOMS same-address service order == lane order: %s (%u violations in %lld trials)\n", hv ? "NO" : "yes", hv, 148LL * 4 * 256 * 20000);
    return 0;
}
