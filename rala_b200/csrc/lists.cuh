// lists.cuh — SoA overlap lists on the device and the tile machinery for ordered compaction.
#pragma once

#include "common.cuh"

namespace rb {

// Device-resident overlap list (the reference's std::vector<std::unique_ptr<Overlap>>), SoA so that
// every per-thread gather / compacted scatter is a 4-byte column access.  b holds b_id | orientation << 31.
struct List {
    uint32_t *a, *b, *ab, *ae, *bb, *be;
    uint8_t* tag;   // Overlap::type() of the entry as of the last pass that computed it
};

struct Entry {
    uint32_t a, b, ori;
    Coords c;
};

__device__ __forceinline__ Entry load_entry(const List& l, uint32_t i) {
    Entry e;
    e.a = l.a[i];
    uint32_t b = l.b[i];
    e.b = b & 0x7FFFFFFFu;
    e.ori = b >> 31;
    e.c.ab = l.ab[i];
    e.c.ae = l.ae[i];
    e.c.bb = l.bb[i];
    e.c.be = l.be[i];
    return e;
}

__device__ __forceinline__ void store_entry(const List& l, uint32_t i, const Entry& e, uint8_t tag) {
    l.a[i] = e.a;
    l.b[i] = e.b | (e.ori << 31);
    l.ab[i] = e.c.ab;
    l.ae[i] = e.c.ae;
    l.bb[i] = e.c.bb;
    l.be[i] = e.c.be;
    l.tag[i] = tag;
}

// Containment events (victim, container, time): SURVEY.md A.3
struct Events {
    uint32_t *v, *c, *t;
};

constexpr int kTileThreads = 256;
constexpr int kTileItems = 4;
constexpr int kTile = kTileThreads * kTileItems;   // entries per tile; item r of thread t is entry r * 256 + t
constexpr int kTileWarps = kTileThreads / 32;

struct TileShared {
    uint32_t cnt_a[kTileItems * kTileWarps];   // index r * 8 + warp : exactly one warp's worth of partials
    uint32_t cnt_b[kTileItems * kTileWarps];
    uint32_t base_a, base_b;
    uint32_t tile;
};

// Ordered positions for up to two output streams.  dest[r] in {0 drop, 1 stream A, 2 stream B}.
// On return pos[r] is the global rank of item r inside its stream (exclusive prefix over all
// earlier entries in entry order).  Returns the packed inclusive totals through *inclusive
// (meaningful in warp 0 for every tile; the last tile's value is the grand total).  Contains two
// __syncthreads.  ITEMS * 8 warps <= 32 partial counts: one warp scans them.
template <int ITEMS>
__device__ __forceinline__ void tile_rank(TileShared& sh, unsigned long long* status, uint32_t tile,
                                          const int (&dest)[ITEMS], uint32_t (&pos)[ITEMS],
                                          unsigned long long* inclusive) {
    static_assert(ITEMS * kTileWarps <= 32, "partials must fit one warp");
    const uint32_t lane = lane_id(), warp = warp_id();
    uint32_t lane_rank[ITEMS];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        uint32_t ma = __ballot_sync(0xFFFFFFFFu, dest[r] == 1);
        uint32_t mb = __ballot_sync(0xFFFFFFFFu, dest[r] == 2);
        uint32_t below = (1u << lane) - 1u;
        lane_rank[r] = dest[r] == 1 ? __popc(ma & below) : __popc(mb & below);
        if (lane == 0) {
            sh.cnt_a[r * kTileWarps + warp] = __popc(ma);
            sh.cnt_b[r * kTileWarps + warp] = __popc(mb);
        }
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t ca = lane < ITEMS * kTileWarps ? sh.cnt_a[lane] : 0u, cb = lane < ITEMS * kTileWarps ? sh.cnt_b[lane] : 0u;
        uint32_t ia = warp_inclusive_scan(ca), ib = warp_inclusive_scan(cb);
        uint32_t ta = __shfl_sync(0xFFFFFFFFu, ia, 31), tb = __shfl_sync(0xFFFFFFFFu, ib, 31);
        unsigned long long agg = pack_counts(ta, tb);
        unsigned long long excl = lookback_exclusive(status, tile, agg);
        if (lane < ITEMS * kTileWarps) {
            sh.cnt_a[lane] = ia - ca;
            sh.cnt_b[lane] = ib - cb;
        }
        if (lane == 0) {
            sh.base_a = count_a(excl);
            sh.base_b = count_b(excl);
        }
        if (inclusive) *inclusive = excl + agg;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        uint32_t base = dest[r] == 1 ? sh.base_a + sh.cnt_a[r * kTileWarps + warp]
                                     : sh.base_b + sh.cnt_b[r * kTileWarps + warp];
        pos[r] = base + lane_rank[r];
    }
}

}  // namespace rb
