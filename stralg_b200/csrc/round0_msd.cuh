// round0_msd.cuh -- interface of the bucketed initial sort (round0_msd.cu) used by sa_build.cu.
#pragma once
#include "engine.h"

namespace b200sa {

// 64 bits of the packed text starting at symbol sym_index (big-endian inside each word)
__device__ __forceinline__ u64 window_at(const u64 *__restrict__ packed, u64 sym_index, int bits) {
    u64 bitpos = sym_index * (u64)bits;
    u64 wi = bitpos >> 6;
    unsigned o = (unsigned)(bitpos & 63);
    u64 hi = packed[wi];
    if (o == 0) return hi;
    u64 lo = packed[wi + 1];
    return (hi << o) | (lo >> (64 - o));
}

// Dense round-0 keys (alphabets that do not fill their symbol width, e.g. DNA + N in 4 bits or the
// 20 amino acids in 8): the key of a suffix is the NUMBER its first K symbols spell in base nsym
// instead of the K * bits raw bits, so that keys spread evenly over their KB bits whatever the
// alphabet -- the bucket tables and the interpolation of the in-SM sort rely on that.  The number is
// formed as hi * powlo + lo from two 32-bit halves of Khi and Klo symbols; K symbols may take up to 128 bits of text.
struct DenseKey {
    u32 nsym;          // 0: raw keys
    u32 Khi, Klo;      // Khi + Klo = K
    u32 powlo;         // nsym ^ Klo
    u64 ptop;          // nsym ^ (K - 1): weight of the leading symbol (keys of consecutive suffixes slide)
};
// the symbol whose first bit is bit x of the 192-bit window H:L:M (symbols never straddle words)
template <int BITS>
__device__ __forceinline__ u32 sym_at192(u64 H, u64 L, u64 M, u32 x) {
    const u32 sel = x >> 6;
    const u64 w = sel == 0 ? H : sel == 1 ? L : M;
    return (u32)(w >> (64 - BITS - (x & 63u))) & ((1u << BITS) - 1u);
}
// key of the next suffix from the key of this one: drop the leading symbol, take in the one after the last
__device__ __forceinline__ u64 dense_key_slide(u64 key, u32 d_out, u32 d_in, const DenseKey &dk) {
    return (key - (u64)d_out * dk.ptop) * dk.nsym + d_in;
}
template <int BITS>
__device__ __forceinline__ u64 dense_key_of(u64 w0, u64 w1, const DenseKey &dk) {  // w0:w1 = 128-bit window starting at the suffix
    u32 hi = 0, lo = 0;
    for (u32 i = 0; i < dk.Khi; ++i) {
        hi = hi * dk.nsym + (u32)(w0 >> (64 - BITS));
        w0 = (w0 << BITS) | (w1 >> (64 - BITS));
        w1 <<= BITS;
    }
    for (u32 i = 0; i < dk.Klo; ++i) {
        lo = lo * dk.nsym + (u32)(w0 >> (64 - BITS));
        w0 = (w0 << BITS) | (w1 >> (64 - BITS));
        w1 <<= BITS;
    }
    return (u64)hi * dk.powlo + lo;
}
// the dense key of the suffix that starts at symbol t (the packed text is padded with zero words)
__device__ __forceinline__ u64 dense_key_at(const u64 *__restrict__ packed, u64 t, int bits, const DenseKey &dk) {
    const u64 w0 = window_at(packed, t, bits), w1 = window_at(packed, t + (u64)(64 / bits), bits);
    return bits == 2 ? dense_key_of<2>(w0, w1, dk) : bits == 4 ? dense_key_of<4>(w0, w1, dk) : dense_key_of<8>(w0, w1, dk);
}

struct MsdPlan {
    int nlevels;       // 1..3 partition levels
    int D[3];          // digit bits per level (multiples of the symbol width; any width with dense keys)
    int BB;            // total bucket bits = D[0] + D[1] + D[2]
    int K;             // symbols in the round-0 key
    int KB;            // key bits: K * bits, or the bits of nsym^K - 1 with dense keys
    DenseKey dense;    // nsym == 0: raw keys
    int pb;            // bits of the preceding symbol carried in an element (0: BWT is gathered afterwards)
    int R;             // key bits left for the in-SM sort = KB - BB
    int dmax;          // largest digit a partition level takes (a multiple of the symbol width)
};

struct Round0Msd {
    // workspace handed in by the caller
    u64 *bufA, *bufB;  // len elements each
    u32 *actbits;      // (len + 31) / 32 + 2 words: bit per ROW whose suffix shares its key ("active")
    u32 *d_primary;    // 1 word
    // results
    MsdPlan plan;
    const u32 *bucket_start;  // 2^BB + 1 entries (workspace arena)
    const u32 *grow;          // [len] by row: group head of the active rows (lives in bufA or bufB)
    u32 *short_rank;          // [1 + 2 * 32] count, then (suffix, final row) of the short suffixes of oversize buckets:
                              // their ranks must be materialised by the caller (workspace arena)
    u32 depth0;               // symbols every active group is known to share: K, or BB / bits when oversize
                              // buckets were emitted unsorted as shallow groups
    u32 levels_added;         // partition levels added to the plan because of a skewed bucket histogram
    u32 shallow_buckets;      // oversize buckets emitted as shallow groups
    u64 shallow_elems;        // suffixes in them
    bool bwt_written;         // BWT rows were emitted by the sort (else gather them from the final SA)
};

// Plans the levels for this text; false when the bucketed sort does not apply (e.g. disabled).
// sym_counts (optional): occurrences of every code, as pack_text counted them.
bool msd_make_plan(u32 len, u32 sigma, int bits, MsdPlan &plan, const u64 *sym_counts = nullptr);

// Sorts all suffixes by their first K symbols.  Buckets that exceed what one SM sorts in shared
// memory are emitted unsorted as shallow groups (depth0 < K); returns false (nothing usable written)
// only when such buckets hold more than 1/8 of the text: the caller then runs the LSD path.
bool round0_msd(DeviceIndex &ix, bool want_bwt, Round0Msd &r);

// not materialised: the rank of the suffix is its row in the round-0 order
static constexpr u32 RANK_NONE = 0xffffffffu;

// rank[sa[g]] = g for every row whose active bit is clear (dense doubling rounds)
void fill_singleton_ranks(const DeviceIndex &ix, const u32 *actbits, u32 *rank);

}  // namespace b200sa
