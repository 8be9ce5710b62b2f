// synth.cu -- deterministic synthetic texts and reads generated on the device (bench/test inputs).
// Counter-based (position -> splitmix64 hash), so the CPU oracle can regenerate the same bytes
// (oracle/stralg_oracle.c: oracle_synth_codes / oracle_synth_reads).
#include "engine.h"

namespace b200sa {

__host__ __device__ __forceinline__ u64 splitmix64(u64 x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) synth_codes_kernel(u8 *__restrict__ text, u64 n, u32 nsym, u64 seed) {
    u64 i0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 > n) return;
    __align__(16) u8 b[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        u64 i = i0 + q;
        b[q] = i < n ? (u8)(1 + (splitmix64(seed + i) >> 33) % nsym) : 0;
    }
    if ((((uintptr_t)text) & 15) == 0 && i0 + 16 <= n + 1) {
        *(uint4 *)(text + i0) = *(uint4 *)b;
    } else {
        for (int q = 0; q < 16; ++q)
            if (i0 + q <= n) text[i0 + q] = b[q];  // text[n] = 0 (sentinel)
    }
}

// read q: with probability miss_per_1024/1024 uniform random symbols, otherwise a copy of
// text[start .. start+m) at a hashed start.  One thread per symbol.
__global__ void __launch_bounds__(256) synth_reads_kernel(const u8 *__restrict__ text, u64 n, u32 nsym,
                                                          u8 *__restrict__ reads, u64 nreads, u32 m,
                                                          u32 miss_per_1024, u64 seed) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nreads * (u64)m) return;
    u64 q = t / m;
    u32 j = (u32)(t % m);
    u64 h = splitmix64(seed ^ (q * 0xD1342543DE82EF95ull));
    bool miss = (h & 1023u) < miss_per_1024;
    u8 c;
    if (miss || n < m) {
        c = (u8)(1 + (splitmix64(h + j) >> 33) % nsym);
    } else {
        u64 start = (h >> 10) % (n - m + 1);
        c = text[start + j];
    }
    reads[t] = c;
}

void synth_codes(u8 *d_text, u64 n, u32 nsym, u64 seed, cudaStream_t st) {
    synth_codes_kernel<<<div_up_u(n / 16 + 1, 256), 256, 0, st>>>(d_text, n, nsym, seed);
    KERNEL_CHECK();
}

void synth_reads(const u8 *d_text, u64 n, u32 nsym, u8 *d_reads, u64 nreads, u32 m, u32 miss_per_1024,
                 u64 seed, cudaStream_t st) {
    if (!nreads || !m) return;
    synth_reads_kernel<<<div_up_u(nreads * (u64)m, 256), 256, 0, st>>>(d_text, n, nsym, d_reads, nreads, m,
                                                                       miss_per_1024, seed);
    KERNEL_CHECK();
}

}  // namespace b200sa
