// classify.cu — K1 / K1b / K1c: overlap trimming + classification, ordered containment as a
// death-time fixed point, chimeric-hill counters, and the ordered list compactions between them.
//
// Replaces (reference file:line): Overlap::trim / Overlap::type overlap.cpp:117-259 as driven by
// graph.cpp:443-518 (classify loop), 722-736 and 801-824 (re-trim, promotion of internals),
// 831-877 (final containment), and Pile::check_chimeric_hills pile.cpp:457-469.
#include "kernels.h"
#include "lists.cuh"

namespace rb {

// =============================================================================================
// K1 (graph.cpp:443-518) runs as two streaming kernels over the 28-byte AoS records the host
// marshals (rala_ovl_t), with the containment fixed point between them:
//   k_classify_events    every record: static gates, trim, type; emits only what the ORDER-DEPENDENT
//                        part needs: the containment events (victim, container, time) of kA/kB records
//                        and the indices of records touching a pile with chimeric hills.
//   k_classify_survivors every record again, once the final pile liveness is known: records whose two
//                        piles are still alive are trimmed, typed and written, in FILE ORDER, to
//                        `overlaps` (kA/kB with a chimeric container, kAB, kBA) or `internals` (kX).
// No intermediate list: on clean data ~95 % of the records are dovetails at classification time but
// only ~7 % survive the dead-pile filter, so materialising the "potential survivors" costs more than
// re-reading the records.
// Both kernels stage tiles of 512 records (14 KiB, contiguous) into shared memory with TMA bulk
// copies (cp.async.bulk + mbarrier), double buffered, and read them back with a 7-word stride
// (odd => bank-conflict free).
// =============================================================================================
constexpr int kRecItems = 2;
constexpr int kRecTile = kTileThreads * kRecItems;       // 512 records
constexpr uint32_t kRecTileWords = kRecTile * 7;

struct RecStage {
    __align__(128) uint32_t rec[2][kRecTileWords];
    __align__(8) uint64_t bar[2];
};

// issue the TMA load of `tile` into buffer b (one thread)
__device__ __forceinline__ void issue_tile(RecStage& st, int b, const uint32_t* __restrict__ rec, uint32_t n, uint32_t tile) {
    const uint32_t base = tile * kRecTile;
    const uint32_t cnt = min((uint32_t) kRecTile, n - base);
    const uint32_t bytes = (cnt * 28u + 15u) & ~15u;   // the record buffer is padded by 16 B past its end
    fence_proxy_async();                               // earlier generic-proxy reads of this buffer come first
    mbar_expect_tx(&st.bar[b], bytes);
    bulk_g2s(st.rec[b], rec + (size_t) base * 7, bytes, &st.bar[b]);
}

struct RecFields {
    Entry e;
    uint32_t flags;
};

__device__ __forceinline__ RecFields read_record(const uint32_t* q) {
    RecFields r;
    r.e.a = q[0];
    r.e.b = q[1];
    r.e.c.ab = q[2];
    r.e.c.ae = q[3];
    r.e.c.bb = q[4];
    r.e.c.be = q[5];
    r.flags = q[6];
    r.e.ori = r.flags & 1u;
    return r;
}

constexpr int kEvStage = 64;   // per-warp staging of events before one aggregated global append

__global__ void __launch_bounds__(kTileThreads) k_classify_events(
    const uint32_t* __restrict__ rec, uint32_t n, uint32_t t0, const uint2* __restrict__ piles, uint32_t n_piles,
    Events ev, uint32_t ev_cap, uint32_t* __restrict__ vcount, uint32_t* __restrict__ hill_rec, uint32_t hill_cap,
    uint32_t* __restrict__ counters) {
    __shared__ RecStage st;
    __shared__ uint32_t s_ev[kTileWarps][3][kEvStage];
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint32_t num_tiles = (n + kRecTile - 1) / kRecTile;
    if (tid == 0) {
        mbar_init(&st.bar[0], 1);
        mbar_init(&st.bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity[2] = {0u, 0u};
    uint32_t staged = 0;   // warp-uniform

    auto flush = [&]() {
        uint32_t gbase = 0;
        if (lane == 0) gbase = atomicAdd(&counters[C_EV], staged);
        gbase = __shfl_sync(0xFFFFFFFFu, gbase, 0);
        for (uint32_t i = lane; i < staged; i += 32) {
            if (gbase + i < ev_cap) {
                ev.v[gbase + i] = s_ev[warp][0][i];
                ev.c[gbase + i] = s_ev[warp][1][i];
                ev.t[gbase + i] = s_ev[warp][2][i];
            }
        }
        __syncwarp();
        staged = 0;
    };

    uint32_t tile = blockIdx.x;
    if (tid == 0 && tile < num_tiles) issue_tile(st, 0, rec, n, tile);
    for (uint32_t it = 0; tile < num_tiles; ++it, tile += gridDim.x) {
        const int b = it & 1;
        if (tid == 0 && tile + gridDim.x < num_tiles) issue_tile(st, b ^ 1, rec, n, tile + gridDim.x);
        mbar_wait(&st.bar[b], parity[b]);
        parity[b] ^= 1u;
        const uint32_t base = tile * kRecTile;
        const uint32_t cnt = min((uint32_t) kRecTile, n - base);
#pragma unroll
        for (int r = 0; r < kRecItems; ++r) {
            const uint32_t idx = r * kTileThreads + tid;
            bool is_ev = false;
            uint32_t evv = 0, evc = 0;
            if (idx < cnt) {
                RecFields f = read_record(st.rec[b] + idx * 7);
                if (!(f.flags & 2u) && f.e.a < n_piles && f.e.b < n_piles) {                 // graph.cpp:450-451
                    const Pile pa = load_pile(piles, f.e.a), pb = load_pile(piles, f.e.b);
                    if (pa.alive() && pb.alive() && trim(f.e.c, f.e.ori, pa, pb)) {           // :451-452
                        if ((pa.flags | pb.flags) & 1u) {                                    // :457-462, resolved by k_hill_coverage
                            uint32_t slot = atomicAdd(&counters[C_HILL], 1u);
                            if (slot < hill_cap) hill_rec[slot] = base + idx;
                        }
                        const uint8_t t = classify(f.e.c, relative(f.e.c, f.e.ori, pa, pb));
                        if (t == kB && !(pb.flags & 2u)) {                                   // :469-474
                            is_ev = true; evv = f.e.a; evc = f.e.b;
                        } else if (t == kA && !(pa.flags & 2u)) {                            // :475-480
                            is_ev = true; evv = f.e.b; evc = f.e.a;
                        }
                    }
                }
            }
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_ev);
            if (m) {
                if (is_ev) {
                    atomicAdd(&vcount[evv], 1u);   // per-victim histogram for the resolution's counting sort
                    const uint32_t p = staged + __popc(m & ((1u << lane) - 1u));
                    s_ev[warp][0][p] = evv;
                    s_ev[warp][1][p] = evc;
                    s_ev[warp][2][p] = t0 + base + idx;
                }
                staged += __popc(m);
                __syncwarp();
                if (staged > kEvStage - 32) flush();
            }
        }
        __syncthreads();   // everyone is done with buffer b before it is refilled
    }
    if (staged) flush();
}

// Survivors pass.  Tiles are independent: a tile appends its survivors (in record order) to scratch
// lists at a base claimed with one atomic, and records (base, count) per tile; k_tile_offsets turns the
// counts into file-order offsets and k_relocate moves each tile's run to its final place.  (A single-pass
// look-back compaction was measured 2.5x slower here: a tile's aggregate is only known after its TMA load
// and two dependent pile gathers, so the look-back chain serialised the waves.)
__global__ void __launch_bounds__(kTileThreads) k_classify_survivors(
    const uint32_t* __restrict__ rec, uint32_t n, const uint2* __restrict__ piles, uint32_t n_piles,
    List tmp_ovl, List tmp_inl, uint32_t cap, TileRuns runs, uint32_t* __restrict__ tmp_counts) {
    __shared__ RecStage st;
    // scratch indexed by the iteration's parity: two barriers per tile suffice (see the end of the loop)
    __shared__ uint32_t s_cnt_a2[2][kRecItems * kTileWarps], s_cnt_b2[2][kRecItems * kTileWarps];
    __shared__ uint32_t s_base_a2[2], s_base_b2[2];
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint32_t num_tiles = (n + kRecTile - 1) / kRecTile;
    if (tid == 0) {
        mbar_init(&st.bar[0], 1);
        mbar_init(&st.bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity[2] = {0u, 0u};
    uint32_t tile = blockIdx.x;
    if (tid == 0 && tile < num_tiles) issue_tile(st, 0, rec, n, tile);
    for (uint32_t it = 0; tile < num_tiles; ++it, tile += gridDim.x) {
        const int b = it & 1;
        uint32_t* s_cnt_a = s_cnt_a2[b];
        uint32_t* s_cnt_b = s_cnt_b2[b];
        uint32_t& s_base_a = s_base_a2[b];
        uint32_t& s_base_b = s_base_b2[b];
        if (tid == 0 && tile + gridDim.x < num_tiles) issue_tile(st, b ^ 1, rec, n, tile + gridDim.x);
        mbar_wait(&st.bar[b], parity[b]);
        parity[b] ^= 1u;
        const uint32_t base = tile * kRecTile;
        const uint32_t cnt = min((uint32_t) kRecTile, n - base);
        int dest[kRecItems];
        Entry e[kRecItems];
        uint8_t tag[kRecItems];
        uint32_t lrank[kRecItems];
#pragma unroll
        for (int r = 0; r < kRecItems; ++r) {
            const uint32_t idx = r * kTileThreads + tid;
            dest[r] = 0;
            tag[r] = kRejected;
            if (idx < cnt) {
                RecFields f = read_record(st.rec[b] + idx * 7);
                if (!(f.flags & 2u) && f.e.a < n_piles && f.e.b < n_piles) {
                    // the table already carries the kills: both piles alive at the end (graph.cpp:493-515)
                    const Pile pa = load_pile(piles, f.e.a), pb = load_pile(piles, f.e.b);   // independent gathers
                    if (pa.alive() && pb.alive() && trim(f.e.c, f.e.ori, pa, pb)) {
                        tag[r] = classify(f.e.c, relative(f.e.c, f.e.ori, pa, pb));
                        dest[r] = tag[r] == kX ? 2 : 1;   // a surviving kA/kB has a chimeric container (:470, :476)
                        e[r] = f.e;
                    }
                }
            }
            const uint32_t ma = __ballot_sync(0xFFFFFFFFu, dest[r] == 1), mb = __ballot_sync(0xFFFFFFFFu, dest[r] == 2);
            const uint32_t below = (1u << lane) - 1u;
            lrank[r] = dest[r] == 1 ? __popc(ma & below) : __popc(mb & below);
            if (lane == 0) {
                s_cnt_a[r * kTileWarps + warp] = __popc(ma);
                s_cnt_b[r * kTileWarps + warp] = __popc(mb);
            }
        }
        __syncthreads();   // counts visible; everyone is done reading buffer b
        if (warp == 0) {
            const uint32_t ca = lane < kRecItems * kTileWarps ? s_cnt_a[lane] : 0u, cb = lane < kRecItems * kTileWarps ? s_cnt_b[lane] : 0u;
            const uint32_t ia = warp_inclusive_scan(ca), ib = warp_inclusive_scan(cb);
            const uint32_t ta = __shfl_sync(0xFFFFFFFFu, ia, 31), tb = __shfl_sync(0xFFFFFFFFu, ib, 31);
            if (lane < kRecItems * kTileWarps) {
                s_cnt_a[lane] = ia - ca;
                s_cnt_b[lane] = ib - cb;
            }
            if (lane == 0) {
                const uint32_t ga = ta ? atomicAdd(&tmp_counts[0], ta) : 0u, gb = tb ? atomicAdd(&tmp_counts[1], tb) : 0u;
                s_base_a = ga;
                s_base_b = gb;
                runs.base_a[tile] = ga;
                runs.cnt_a[tile] = ta;
                runs.base_b[tile] = gb;
                runs.cnt_b[tile] = tb;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kRecItems; ++r) {
            if (dest[r] == 1) {
                const uint32_t p = s_base_a + s_cnt_a[r * kTileWarps + warp] + lrank[r];
                if (p < cap) store_entry(tmp_ovl, p, e[r], tag[r]);
            } else if (dest[r] == 2) {
                const uint32_t p = s_base_b + s_cnt_b[r * kTileWarps + warp] + lrank[r];
                if (p < cap) store_entry(tmp_inl, p, e[r], tag[r]);
            }
        }
        // no trailing barrier: this parity's scratch is next written two iterations from now (two barriers away),
        // and staging buffer b is refilled by the TMA thread 0 issues at the top of the next iteration, i.e. after
        // it passed this iteration's second barrier, which every thread reaches only after reading buffer b
    }
}

// scratch run -> final position, one thread per output entry: the tile of entry j is found by binary
// search in the scanned offsets (off[t] <= j < off[t+1]); writes are fully coalesced
__global__ void k_relocate(List tmp, List out, uint32_t cap, const uint32_t* __restrict__ base, const uint32_t* __restrict__ off,
                           uint32_t num_tiles) {
    const uint32_t total = min(off[num_tiles], cap);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = num_tiles;   // largest t with off[t] <= j
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (off[mid] <= j) lo = mid; else hi = mid;
        }
        const uint32_t src = base[lo] + (j - off[lo]);
        if (src < cap) {
            out.a[j] = tmp.a[src]; out.b[j] = tmp.b[src];
            out.ab[j] = tmp.ab[src]; out.ae[j] = tmp.ae[src];
            out.bb[j] = tmp.bb[src]; out.be[j] = tmp.be[src];
            out.tag[j] = tmp.tag[src];
        }
    }
}

// K1c: Pile::check_chimeric_hills (pile.cpp:457-469) for every PROCESSED record that touches a pile
// with hills.  Processed = passed the static gates and both piles were still alive at its time.
__global__ void k_hill_coverage(const uint32_t* __restrict__ rec, uint32_t t0, const uint2* __restrict__ piles,
                                const uint32_t* __restrict__ hill_rec, uint32_t hill_cap,
                                const uint32_t* __restrict__ hill_pile, const uint32_t* __restrict__ hill_begin,
                                const uint32_t* __restrict__ hill_end, uint32_t n_hills, uint32_t* __restrict__ hill_cov,
                                const uint32_t* __restrict__ dbuf, uint32_t n_piles, const uint32_t* __restrict__ counters) {
    const uint32_t n = min(counters[C_HILL], hill_cap);
    const uint32_t* D = dbuf + (size_t) counters[C_DSEL] * n_piles;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t idx = hill_rec[i];
        const uint32_t* q = rec + (size_t) idx * 7;
        Entry e;
        e.a = q[0]; e.b = q[1]; e.c.ab = q[2]; e.c.ae = q[3]; e.c.bb = q[4]; e.c.be = q[5]; e.ori = q[6] & 1u;
        const uint32_t t = t0 + idx;
        if (D[e.a] < t || D[e.b] < t) continue;   // a pile was already dead: transmute() rejected it (overlap.cpp:51,70)
        const Pile pa = load_pile(piles, e.a), pb = load_pile(piles, e.b);
        trim(e.c, e.ori, pa, pb);
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const Pile& p = side ? pb : pa;
            if (!(p.flags & 1u)) continue;
            const uint32_t id = side ? e.b : e.a;
            const uint32_t lo = p.begin + (side ? e.c.bb : e.c.ab);   // begin_ added to absolute coordinates, as in :459-462
            const uint32_t hi = p.begin + (side ? e.c.be : e.c.ae);
            uint32_t l = 0, h = n_hills;
            while (l < h) {
                uint32_t m = (l + h) >> 1;
                if (hill_pile[m] < id) l = m + 1; else h = m;
            }
            for (uint32_t k = l; k < n_hills && hill_pile[k] == id; ++k) {
                if (lo < hill_begin[k] && hi > hill_end[k]) atomicAdd(&hill_cov[k], 1u);
            }
        }
    }
}

// Piles with a finite death time die (piles_[x].reset(), graph.cpp:471,477,838,842).
__global__ void k_apply_deaths(uint2* __restrict__ piles, const uint32_t* __restrict__ dbuf, uint32_t n_piles,
                               const uint32_t* __restrict__ counters) {
    const uint32_t* D = dbuf + (size_t) counters[C_DSEL] * n_piles;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_piles; i += gridDim.x * blockDim.x) {
        if (D[i] != kInf) piles[i] = make_uint2(0u, 0u);
    }
}

// =============================================================================================
// Ordered list passes.  MODE selects what happens to an entry:
//   kSplitAlive   P -> overlaps (tag != kX) / internals (tag == kX), both piles alive   graph.cpp:493-515
//   kRetrim       list -> list, dropping entries trim() rejects                          :722-736, 801-807
//   kPromote      internals -> internals / appended to overlaps when now a dovetail      :809-824
//   kFinalOvl     overlaps -> overlaps, keeping entries with both piles alive and not kA/kB  :831-848, 869-877
//   kFinalInt     internals -> internals (alive at their own time, not kA/kB)            :849-867
// =============================================================================================
enum ListMode { kSplitAlive = 0, kRetrim = 1, kPromote = 2, kFinalOvl = 3, kFinalInt = 4 };

template <int MODE>
__global__ void __launch_bounds__(kTileThreads) k_list_pass(
    List in, const uint32_t* __restrict__ n_in_ptr, uint32_t in_cap, const uint2* __restrict__ piles,
    List out_a, uint32_t* __restrict__ n_out_a, List out_b, uint32_t* __restrict__ n_out_b,
    const uint32_t* __restrict__ b_base_ptr, uint32_t cap, const uint32_t* __restrict__ dbuf, uint32_t n_piles, const uint32_t* __restrict__ time_base_ptr,
    const uint32_t* __restrict__ counters, unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket) {
    __shared__ TileShared sh;
    const uint32_t tid = threadIdx.x;
    const uint32_t n = min(*n_in_ptr, in_cap);
    const uint32_t num_tiles = (n + kTile - 1) / kTile;
    // kPromote appends behind the current end of `overlaps` (count in a slot this kernel never writes)
    const uint32_t b_off = b_base_ptr ? *b_base_ptr : 0u;
    const uint32_t* D = dbuf ? dbuf + (size_t) counters[C_DSEL] * n_piles : nullptr;
    const uint32_t time_base = time_base_ptr ? *time_base_ptr : 0u;
    __syncthreads();

    while (true) {
        if (tid == 0) sh.tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = sh.tile;
        if (tile >= num_tiles) break;
        const uint32_t base = tile * kTile;

        Entry e[kTileItems];
        uint8_t tag[kTileItems];
        int dest[kTileItems];
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) {
            const uint32_t idx = base + r * kTileThreads + tid;
            dest[r] = 0;
            tag[r] = kRejected;
            if (idx < n) {
                e[r] = load_entry(in, idx);
                tag[r] = in.tag[idx];
                const Pile pa = load_pile(piles, e[r].a), pb = load_pile(piles, e[r].b);
                if (MODE == kSplitAlive) {
                    if (pa.alive() && pb.alive()) dest[r] = tag[r] == kX ? 2 : 1;
                } else if (MODE == kRetrim) {
                    if (pa.alive() && pb.alive() && trim(e[r].c, e[r].ori, pa, pb)) dest[r] = 1;
                } else if (MODE == kPromote) {
                    if (pa.alive() && pb.alive() && trim(e[r].c, e[r].ori, pa, pb)) {
                        tag[r] = classify(e[r].c, relative(e[r].c, e[r].ori, pa, pb));
                        dest[r] = (tag[r] == kAB || tag[r] == kBA) ? 2 : 1;
                    }
                } else if (MODE == kFinalOvl) {
                    if (pa.alive() && pb.alive() && tag[r] != kA && tag[r] != kB) dest[r] = 1;
                } else {   // kFinalInt: alive before the pass (alive now, or killed by it) and still alive at its own time
                    const uint32_t t = time_base + idx;
                    const uint32_t da = D[e[r].a], db = D[e[r].b];
                    if ((pa.alive() || da != kInf) && (pb.alive() || db != kInf) && da > t && db > t && tag[r] != kA &&
                        tag[r] != kB)
                        dest[r] = 1;
                }
            }
        }
        uint32_t pos[kTileItems];
        unsigned long long inclusive = 0;
        tile_rank<kTileItems>(sh, status, tile, dest, pos, &inclusive);
        if (tid == 0 && tile == num_tiles - 1) {
            *n_out_a = count_a(inclusive);
            if (n_out_b) *n_out_b = b_off + count_b(inclusive);
        }
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) {
            if (dest[r] == 1 && pos[r] < cap) store_entry(out_a, pos[r], e[r], tag[r]);
            if (dest[r] == 2 && b_off + pos[r] < cap) store_entry(out_b, b_off + pos[r], e[r], tag[r]);
        }
        __syncthreads();
    }
}

// Final containment, classification half (graph.cpp:831-866): type of every entry of `overlaps`
// then `internals` against the final pile table; kA/kB become events with time = position in the
// concatenation.  No chimeric gating in this pass.
__global__ void k_classify_final(List lst, const uint32_t* __restrict__ n_ptr, uint32_t cap,
                                 const uint32_t* __restrict__ time_base_ptr, const uint2* __restrict__ piles, Events ev,
                                 uint32_t ev_cap, uint32_t* __restrict__ vcount, uint32_t* __restrict__ counters) {
    const uint32_t n = min(*n_ptr, cap);
    const uint32_t time_base = time_base_ptr ? *time_base_ptr : 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Entry e = load_entry(lst, i);
        const Pile pa = load_pile(piles, e.a), pb = load_pile(piles, e.b);
        uint8_t t = kRejected;
        if (pa.alive() && pb.alive()) {
            t = classify(e.c, relative(e.c, e.ori, pa, pb));
            if (t == kA || t == kB) {
                // warp-aggregated append
                uint32_t m = __activemask();
                m = __match_any_sync(m, 1);
                uint32_t leader = __ffs(m) - 1, rank = __popc(m & ((1u << lane_id()) - 1u));
                uint32_t gbase = 0;
                if (lane_id() == leader) gbase = atomicAdd(&counters[C_EV], (uint32_t) __popc(m));
                gbase = __shfl_sync(m, gbase, leader);
                uint32_t p = gbase + rank;
                if (p < ev_cap) {
                    ev.v[p] = t == kA ? e.b : e.a;
                    ev.c[p] = t == kA ? e.a : e.b;
                    ev.t[p] = time_base + i;
                }
                atomicAdd(&vcount[t == kA ? e.b : e.a], 1u);
            }
        }
        lst.tag[i] = t;
    }
}

// Stateless unit stage: trim + type for n independent records (C ABI rala_b200_trim_classify).
__global__ void k_trim_classify_aos(uint32_t* __restrict__ rec, uint32_t n, const uint2* __restrict__ piles,
                                    uint32_t n_piles, uint8_t* __restrict__ type_out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t* q = rec + (size_t) i * 7;
        Entry e;
        e.a = q[0]; e.b = q[1]; e.c.ab = q[2]; e.c.ae = q[3]; e.c.bb = q[4]; e.c.be = q[5]; e.ori = q[6] & 1u;
        uint8_t t = kRejected;
        if (!(q[6] & 2u) && e.a < n_piles && e.b < n_piles) {
            const Pile pa = load_pile(piles, e.a), pb = load_pile(piles, e.b);
            if (pa.alive() && pb.alive() && trim(e.c, e.ori, pa, pb)) {
                t = classify(e.c, relative(e.c, e.ori, pa, pb));
                q[2] = e.c.ab; q[3] = e.c.ae; q[4] = e.c.bb; q[5] = e.c.be;
            }
        }
        type_out[i] = t;
    }
}

__global__ void k_fill_u32(uint32_t* __restrict__ p, uint32_t v, size_t n) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}

// host table {begin,end} + flag bytes -> packed device table
__global__ void k_pack_piles(const uint2* __restrict__ in, const uint8_t* __restrict__ flags, uint2* __restrict__ out,
                             uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint2 v = in[i];
        uint32_t f = flags ? (flags[i] & 3u) : 0u;
        out[i] = v.y == 0u ? make_uint2(0u, 0u) : make_uint2(v.x, (v.y & kEndMask) | (f << 30));
    }
}

__global__ void k_unpack_piles(const uint2* __restrict__ in, uint2* __restrict__ out, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint2 v = in[i];
        out[i] = make_uint2(v.x, v.y & kEndMask);
    }
}

// SoA list -> rala_ovl_t rows (download path)
__global__ void k_list_to_aos(List l, const uint32_t* __restrict__ n_ptr, uint32_t cap, uint32_t* __restrict__ out) {
    const uint32_t n = min(*n_ptr, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Entry e = load_entry(l, i);
        uint32_t* q = out + (size_t) i * 7;
        q[0] = e.a; q[1] = e.b; q[2] = e.c.ab; q[3] = e.c.ae; q[4] = e.c.bb; q[5] = e.c.be; q[6] = e.ori;
    }
}

__global__ void k_list_connections(List l, const uint32_t* __restrict__ n_ptr, uint32_t cap, uint32_t* __restrict__ out) {
    const uint32_t n = min(*n_ptr, cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        out[2 * (size_t) i] = l.a[i];
        out[2 * (size_t) i + 1] = l.b[i] & 0x7FFFFFFFu;
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
static inline int grid_for(uint64_t n, int per_block, int max_blocks) {
    uint64_t b = (n + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return (int) (b < (uint64_t) max_blocks ? b : (uint64_t) max_blocks);
}

void launch_classify_events(Launch& L, const uint32_t* rec, uint32_t n, uint32_t t0, const uint2* piles, uint32_t n_piles,
                            Events ev, uint32_t ev_cap, uint32_t* vcount, uint32_t* hill_rec, uint32_t hill_cap, uint32_t* counters) {
    if (n == 0) return;
    int grid = grid_for(n, kRecTile, kNumSMs * 6);
    k_classify_events<<<grid, kTileThreads, 0, L.stream>>>(rec, n, t0, piles, n_piles, ev, ev_cap, vcount, hill_rec, hill_cap, counters);
    L.count++;
}

void launch_classify_survivors(Launch& L, const uint32_t* rec, uint32_t n, const uint2* piles, uint32_t n_piles, List tmp_ovl,
                               List tmp_inl, List ovl, uint32_t* n_ovl, List inl, uint32_t* n_inl, uint32_t cap, TileRuns runs,
                               uint32_t* tmp_counts, unsigned long long* status[2], uint32_t* ticket[2]) {
    if (n == 0) return;   // n_ovl / n_inl were zeroed by the caller
    const uint32_t num_tiles = (n + kRecTile - 1) / kRecTile;
    // counts of the sentinel tile [num_tiles] must be zero so that the scans end with the totals
    cudaMemsetAsync(runs.cnt_a + num_tiles, 0, 4, L.stream);
    cudaMemsetAsync(runs.cnt_b + num_tiles, 0, 4, L.stream);
    int grid = grid_for(n, kRecTile, kNumSMs * 6);
    k_classify_survivors<<<grid, kTileThreads, 0, L.stream>>>(rec, n, piles, n_piles, tmp_ovl, tmp_inl, cap, runs, tmp_counts);
    L.count++;
    launch_scan_u32(L, runs.cnt_a, runs.off_a, num_tiles + 1, status[0], ticket[0]);
    launch_scan_u32(L, runs.cnt_b, runs.off_b, num_tiles + 1, status[1], ticket[1]);
    cudaMemcpyAsync(n_ovl, runs.off_a + num_tiles, 4, cudaMemcpyDeviceToDevice, L.stream);
    cudaMemcpyAsync(n_inl, runs.off_b + num_tiles, 4, cudaMemcpyDeviceToDevice, L.stream);
    k_relocate<<<grid_for(cap, 256, kNumSMs * 8), 256, 0, L.stream>>>(tmp_ovl, ovl, cap, runs.base_a, runs.off_a, num_tiles);
    k_relocate<<<grid_for(cap, 256, kNumSMs * 8), 256, 0, L.stream>>>(tmp_inl, inl, cap, runs.base_b, runs.off_b, num_tiles);
    L.count += 2;
}

uint32_t classify_num_tiles(uint32_t n) { return (n + kRecTile - 1) / kRecTile; }

void launch_hill_coverage(Launch& L, const uint32_t* rec, uint32_t t0, const uint2* piles, const uint32_t* hill_rec,
                          uint32_t hill_cap, const uint32_t* hill_pile, const uint32_t* hill_begin,
                          const uint32_t* hill_end, uint32_t n_hills, uint32_t* hill_cov, const uint32_t* dbuf,
                          uint32_t n_piles, const uint32_t* counters) {
    k_hill_coverage<<<kNumSMs * 2, 256, 0, L.stream>>>(rec, t0, piles, hill_rec, hill_cap, hill_pile, hill_begin, hill_end,
                                                       n_hills, hill_cov, dbuf, n_piles, counters);
    L.count++;
}

void launch_apply_deaths(Launch& L, uint2* piles, const uint32_t* dbuf, uint32_t n_piles, const uint32_t* counters) {
    if (n_piles == 0) return;
    k_apply_deaths<<<grid_for(n_piles, 256, kNumSMs * 8), 256, 0, L.stream>>>(piles, dbuf, n_piles, counters);
    L.count++;
}

void launch_list_pass(Launch& L, int mode, List in, const uint32_t* n_in, uint32_t in_cap, const uint2* piles, List out_a,
                      uint32_t* n_out_a, List out_b, uint32_t* n_out_b, const uint32_t* b_base, uint32_t cap, const uint32_t* dbuf,
                      uint32_t n_piles, const uint32_t* time_base, const uint32_t* counters, unsigned long long* status,
                      uint32_t* ticket) {
    int grid = grid_for(in_cap, kTile, kNumSMs * 8);
#define RB_LP(M) k_list_pass<M><<<grid, kTileThreads, 0, L.stream>>>(in, n_in, in_cap, piles, out_a, n_out_a, out_b, \
        n_out_b, b_base, cap, dbuf, n_piles, time_base, counters, status, ticket)
    switch (mode) {
        case kSplitAlive: RB_LP(kSplitAlive); break;
        case kRetrim: RB_LP(kRetrim); break;
        case kPromote: RB_LP(kPromote); break;
        case kFinalOvl: RB_LP(kFinalOvl); break;
        default: RB_LP(kFinalInt); break;
    }
#undef RB_LP
    L.count++;
}

void launch_classify_final(Launch& L, List lst, const uint32_t* n_ptr, uint32_t cap, const uint32_t* time_base,
                           const uint2* piles, Events ev, uint32_t ev_cap, uint32_t* vcount, uint32_t* counters) {
    k_classify_final<<<grid_for(cap, 256, kNumSMs * 8), 256, 0, L.stream>>>(lst, n_ptr, cap, time_base, piles, ev, ev_cap,
                                                                            vcount, counters);
    L.count++;
}

void launch_trim_classify_aos(Launch& L, uint32_t* rec, uint32_t n, const uint2* piles, uint32_t n_piles, uint8_t* type_out) {
    if (n == 0) return;
    k_trim_classify_aos<<<grid_for(n, 256, kNumSMs * 8), 256, 0, L.stream>>>(rec, n, piles, n_piles, type_out);
    L.count++;
}

void launch_fill_u32(Launch& L, uint32_t* p, uint32_t v, size_t n) {
    if (n == 0) return;
    k_fill_u32<<<grid_for(n, 1024, kNumSMs * 8), 256, 0, L.stream>>>(p, v, n);
    L.count++;
}

void launch_pack_piles(Launch& L, const uint2* in, const uint8_t* flags, uint2* out, uint32_t n) {
    if (n == 0) return;
    k_pack_piles<<<grid_for(n, 256, kNumSMs * 8), 256, 0, L.stream>>>(in, flags, out, n);
    L.count++;
}

void launch_unpack_piles(Launch& L, const uint2* in, uint2* out, uint32_t n) {
    if (n == 0) return;
    k_unpack_piles<<<grid_for(n, 256, kNumSMs * 8), 256, 0, L.stream>>>(in, out, n);
    L.count++;
}

void launch_list_to_aos(Launch& L, List l, const uint32_t* n_ptr, uint32_t cap, uint32_t* out) {
    k_list_to_aos<<<grid_for(cap, 256, kNumSMs * 8), 256, 0, L.stream>>>(l, n_ptr, cap, out);
    L.count++;
}

void launch_list_connections(Launch& L, List l, const uint32_t* n_ptr, uint32_t cap, uint32_t* out) {
    k_list_connections<<<grid_for(cap, 256, kNumSMs * 8), 256, 0, L.stream>>>(l, n_ptr, cap, out);
    L.count++;
}

}  // namespace rb
