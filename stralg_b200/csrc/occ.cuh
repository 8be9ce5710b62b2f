// occ.cuh -- sampled, popcount-friendly occurrence table: device-side rank queries.
//
// O(a, i) = #{k < i : bwt[k] == a}, i in [0, len]  (stralg/bwt.c:47-65, macro bwt.h:48-50).
// The reference stores the dense (len+1) x sigma matrix; here one block covers 64 BWT rows:
//
//   OCC_DNA32 (sigma <= 5): 32 bytes = one DRAM sector per query
//        u32 cnt[4]   occurrences of codes 1..4 in all rows before the block
//        u64 bits[2]  64 rows x 2 bits, row k of the block at bits [2k, 2k+1] of word k/32,
//                     value = code - 1; the sentinel row is stored as 0 and corrected by `primary`
//   OCC_BYTE (any sigma): hdr_words u32 counts for codes 1..sigma-1 (padded to 4) + 64 row bytes
#pragma once
#include "common.cuh"

namespace b200sa {

struct OccView {
    const u8 *blocks;
    u32 block_bytes;
    u32 hdr_words;   // OCC_BYTE only
    u32 primary;     // row whose BWT symbol is the sentinel
    u32 sigma;
    int layout;      // OccLayout
};

struct __align__(16) DnaBlock {
    u32 cnt[4];
    u64 bits[2];
};

__device__ __forceinline__ u32 dna_match_count(u64 w0, u64 w1, u32 sym, u32 r) {
    // number of rows k < r (r in [0, 64]) whose 2-bit value equals sym
    const u64 lowbits = 0x5555555555555555ull;
    u64 pat = (u64)sym * lowbits;
    u64 y0 = w0 ^ pat, y1 = w1 ^ pat;
    u64 m0 = ~(y0 | (y0 >> 1)) & lowbits;
    u64 m1 = ~(y1 | (y1 >> 1)) & lowbits;
    u32 r0 = r < 32 ? r : 32;
    u32 r1 = r > 32 ? r - 32 : 0;
    u64 k0 = r0 == 32 ? ~0ull : ((1ull << (2 * r0)) - 1ull);
    u64 k1 = r1 == 32 ? ~0ull : ((1ull << (2 * r1)) - 1ull);
    return __popcll(m0 & k0) + __popcll(m1 & k1);
}

__device__ __forceinline__ u32 occ_dna(const OccView &ov, u32 a, u32 i) {
    // a in 1..4
    const DnaBlock *blk = (const DnaBlock *)ov.blocks + (i >> 6);
    uint4 h = *(const uint4 *)blk;                  // cnt[0..3]
    uint4 p = *((const uint4 *)blk + 1);            // bits
    u64 w0 = ((u64)p.y << 32) | p.x, w1 = ((u64)p.w << 32) | p.z;
    u32 base = a == 1 ? h.x : a == 2 ? h.y : a == 3 ? h.z : h.w;
    u32 r = i & 63u;
    u32 c = base + dna_match_count(w0, w1, a - 1, r);
    // the sentinel row was packed as value 0 (code 1): take it back out
    if (a == 1) {
        u32 start = i & ~63u;
        if (ov.primary >= start && ov.primary < i) c -= 1;
    }
    return c;
}

__device__ __forceinline__ u32 byte_match_count(const u8 *rows, u32 a, u32 r) {
    // rows: 64 bytes, 16-byte aligned; count rows k < r equal to a
    const u64 ones = 0x0101010101010101ull, low7 = 0x7f7f7f7f7f7f7f7full;
    const uint4 *q = (const uint4 *)rows;
    u32 c = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        if ((u32)(v * 16) >= r) break;
        uint4 x = q[v];
        u64 w[2] = {((u64)x.y << 32) | x.x, ((u64)x.w << 32) | x.z};
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            u32 first = v * 16 + hh * 8;
            if (first >= r) break;
            u64 y = w[hh] ^ (ones * a);
            u64 z = ~(((y & low7) + low7) | y | low7);  // 0x80 in every byte of y that is zero
            u32 left = r - first;
            if (left < 8) z &= (1ull << (8 * left)) - 1ull;
            c += __popcll(z);
        }
    }
    return c;
}

__device__ __forceinline__ u32 occ_byte(const OccView &ov, u32 a, u32 i) {
    const u8 *blk = ov.blocks + (size_t)(i >> 6) * ov.block_bytes;
    u32 base = ((const u32 *)blk)[a - 1];
    return base + byte_match_count(blk + (size_t)ov.hdr_words * 4, a, i & 63u);
}

// O(a, i) for any a in [0, sigma)
__device__ __forceinline__ u32 occ_any(const OccView &ov, u32 a, u32 i) {
    if (a == 0) return i > ov.primary ? 1u : 0u;
    if (ov.layout == 1) return occ_dna(ov, a, i);
    return occ_byte(ov, a, i);
}

struct DeviceIndex;
OccView occ_view(const DeviceIndex &ix);  // bwt_occ.cu

}  // namespace b200sa
