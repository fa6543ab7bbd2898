// classify.cu — K1 / K1b / K1c: overlap trimming + classification, ordered containment as a
// death-time fixed point, chimeric-hill counters, and the ordered list compactions between them.
//
// Replaces (reference file:line): Overlap::trim / Overlap::type overlap.cpp:117-259 as driven by
// graph.cpp:443-518 (classify loop), 722-736 and 801-824 (re-trim, promotion of internals),
// 831-877 (final containment), and Pile::check_chimeric_hills pile.cpp:457-469.
#include <cooperative_groups.h>

#include "kernels.h"
#include "lists.cuh"

namespace cg = cooperative_groups;

namespace rb {

// =============================================================================================
// K1 first pass (graph.cpp:448-488): one thread per record.  Records arrive as the 28-byte AoS the
// host marshals (rala_ovl_t); a tile of 1024 records (28 KiB, contiguous) is staged into shared
// memory by one TMA bulk copy and read back with a 7-word stride (odd => conflict free).
// Outputs, all in one pass over the records:
//   P      records that can still survive (kX, dovetails, kA/kB whose container is chimeric), trimmed,
//          in FILE ORDER (single-pass decoupled look-back compaction)            -> 25 B each
//   events (victim, container, time) of every kA/kB that may kill                -> 12 B each, any order
//   hills  indices of records that touch a pile with chimeric hills (rare)       ->  4 B each, any order
// =============================================================================================
__global__ void __launch_bounds__(kTileThreads) k_classify_first(
    const uint32_t* __restrict__ rec, uint32_t n, uint32_t t0, const uint2* __restrict__ piles, uint32_t n_piles,
    List P, uint32_t p_cap, Events ev, uint32_t ev_cap, uint32_t* __restrict__ hill_rec, uint32_t hill_cap,
    uint32_t* __restrict__ counters, unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket) {
    __shared__ __align__(128) uint32_t s_rec[kTile * 7];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ TileShared sh;
    __shared__ uint32_t s_ev_cnt[kTileItems * kTileWarps];
    __shared__ uint32_t s_ev_base;

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint32_t num_tiles = (n + kTile - 1) / kTile;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;

    while (true) {
        if (tid == 0) sh.tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = sh.tile;
        if (tile >= num_tiles) break;
        const uint32_t base = tile * kTile;
        const uint32_t cnt = min((uint32_t) kTile, n - base);
        if (cnt == kTile) {
            if (tid == 0) {
                fence_proxy_async();   // earlier generic-proxy reads of s_rec are ordered before the async write
                mbar_expect_tx(&s_bar, kTile * 28u);
                bulk_g2s(s_rec, rec + (size_t) base * 7, kTile * 28u, &s_bar);
            }
            mbar_wait(&s_bar, parity);
            parity ^= 1u;
        } else {
            for (uint32_t i = tid; i < cnt * 7; i += kTileThreads) s_rec[i] = rec[(size_t) base * 7 + i];
            __syncthreads();
        }

        Entry e[kTileItems];
        uint8_t tag[kTileItems];
        int dest[kTileItems];
        uint32_t evv[kTileItems], evc[kTileItems];
        bool is_ev[kTileItems];
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) {
            const uint32_t idx = r * kTileThreads + tid;
            dest[r] = 0;
            is_ev[r] = false;
            tag[r] = kRejected;
            if (idx < cnt) {
                const uint32_t* q = s_rec + idx * 7;
                e[r].a = q[0];
                e[r].b = q[1];
                e[r].c.ab = q[2];
                e[r].c.ae = q[3];
                e[r].c.bb = q[4];
                e[r].c.be = q[5];
                const uint32_t flags = q[6];
                e[r].ori = flags & 1u;
                if (!(flags & 2u) && e[r].a < n_piles && e[r].b < n_piles) {          // graph.cpp:450-451
                    const Pile pa = load_pile(piles, e[r].a), pb = load_pile(piles, e[r].b);
                    if (pa.alive() && pb.alive() && trim(e[r].c, e[r].ori, pa, pb)) {   // :451-452
                        if ((pa.flags | pb.flags) & 1u) {                               // :457-462, resolved later
                            uint32_t slot = atomicAdd(&counters[C_HILL], 1u);
                            if (slot < hill_cap) hill_rec[slot] = base + idx;
                        }
                        const uint8_t t = classify(e[r].c, relative(e[r].c, e[r].ori, pa, pb));
                        tag[r] = t;
                        if (t == kB && !(pb.flags & 2u)) {                              // :469-474
                            is_ev[r] = true; evv[r] = e[r].a; evc[r] = e[r].b;
                        } else if (t == kA && !(pa.flags & 2u)) {                       // :475-480
                            is_ev[r] = true; evv[r] = e[r].b; evc[r] = e[r].a;
                        } else {
                            dest[r] = 1;
                        }
                    }
                }
            }
        }

        // events: unordered, one global atomic per tile
        uint32_t ev_rank[kTileItems];
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) {
            uint32_t m = __ballot_sync(0xFFFFFFFFu, is_ev[r]);
            ev_rank[r] = __popc(m & ((1u << lane) - 1u));
            if (lane == 0) s_ev_cnt[r * kTileWarps + warp] = __popc(m);
        }
        uint32_t pos[kTileItems];
        unsigned long long inclusive = 0;
        tile_rank(sh, status, tile, dest, pos, &inclusive);   // (first __syncthreads inside publishes s_ev_cnt)
        if (warp == 0) {
            uint32_t c = s_ev_cnt[lane];
            uint32_t inc = warp_inclusive_scan(c);
            uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
            uint32_t gbase = 0;
            if (lane == 0 && total) gbase = atomicAdd(&counters[C_EV], total);
            gbase = __shfl_sync(0xFFFFFFFFu, gbase, 0);
            s_ev_cnt[lane] = gbase + inc - c;
            if (lane == 0 && tile == num_tiles - 1) counters[C_P] = count_a(inclusive);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kTileItems; ++r) {
            if (dest[r] == 1 && pos[r] < p_cap) store_entry(P, pos[r], e[r], tag[r]);
            if (is_ev[r]) {
                uint32_t p = s_ev_cnt[r * kTileWarps + warp] + ev_rank[r];
                if (p < ev_cap) {
                    ev.v[p] = evv[r];
                    ev.c[p] = evc[r];
                    ev.t[p] = t0 + base + r * kTileThreads + tid;
                }
            }
        }
        __syncthreads();   // s_rec, sh and s_ev_cnt are reused by the next tile
    }
}

// =============================================================================================
// K1b: ordered containment as a death-time fixed point.
//   D_{r+1}[x] = min{ t_i : victim_i = x and D_r[container_i] > t_i },  D_0 = +inf.
// One cooperative persistent kernel; four rotating D buffers make a round ONE pass over the events
// and ONE grid barrier:  in round r the pass (1) builds D_{r+1} with atomicMin, (2) checks
// D_r == D_{r-1} on the victims (only victims ever change), (3) resets the victims' slots of the
// buffer round r+1 will build.  When the check finds no difference D_r is the answer.
// =============================================================================================
__global__ void __launch_bounds__(256) k_containment_fixpoint(Events ev, const uint32_t* __restrict__ n_events,
                                                             uint32_t ev_cap, uint32_t* __restrict__ dbuf,
                                                             uint32_t n_piles, uint32_t* __restrict__ flags,
                                                             uint32_t* __restrict__ counters) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t n = min(*n_events, ev_cap);
    const uint32_t stride = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t* D[4] = {dbuf, dbuf + n_piles, dbuf + 2 * (size_t) n_piles, dbuf + 3 * (size_t) n_piles};
    // buffers arrive filled with +inf; D[0] = D_0.  flags[r % 3] is the "changed" flag of round r (three slots:
    // slot (r+1)%3 is cleared during round r, after every thread has finished reading it for round r-2).
    uint32_t round = 0;
    uint32_t result = 0;
    if (n == 0) {
        if (gtid == 0) { counters[C_ROUNDS] = 0; counters[C_DSEL] = 0; }
        return;
    }
    while (true) {
        const uint32_t* cur = D[round & 3];                 // D_r
        const uint32_t* prev = D[(round + 3) & 3];          // D_{r-1}
        uint32_t* next = D[(round + 1) & 3];                // D_{r+1}, pre-reset
        uint32_t* next2 = D[(round + 2) & 3];               // buffer of round r+1, reset now
        bool changed = false;
        for (uint32_t i = gtid; i < n; i += stride) {
            const uint32_t v = ev.v[i], c = ev.c[i], t = ev.t[i];
            if (cur[c] > t) atomicMin(&next[v], t);
            if (round > 0 && cur[v] != prev[v]) changed = true;
            next2[v] = kInf;
        }
        if (round > 0 && changed) flags[round % 3] = 1u;
        if (gtid == 0) flags[(round + 1) % 3] = 0u;
        grid.sync();
        if (round > 0 && flags[round % 3] == 0u) {
            result = round & 3;
            break;
        }
        ++round;
    }
    if (gtid == 0) {
        counters[C_ROUNDS] = round;
        counters[C_DSEL] = result;
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
                } else {   // kFinalInt: piles here are the table BEFORE the final kills were applied
                    const uint32_t t = time_base + idx;
                    if (pa.alive() && pb.alive() && D[e[r].a] > t && D[e[r].b] > t && tag[r] != kA && tag[r] != kB)
                        dest[r] = 1;
                }
            }
        }
        uint32_t pos[kTileItems];
        unsigned long long inclusive = 0;
        tile_rank(sh, status, tile, dest, pos, &inclusive);
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
                                 uint32_t ev_cap, uint32_t* __restrict__ counters) {
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

void launch_classify_first(Launch& L, const uint32_t* rec, uint32_t n, uint32_t t0, const uint2* piles, uint32_t n_piles,
                           List P, uint32_t p_cap, Events ev, uint32_t ev_cap, uint32_t* hill_rec, uint32_t hill_cap,
                           uint32_t* counters, unsigned long long* status, uint32_t* ticket) {
    if (n == 0) return;
    int grid = grid_for(n, kTile, kNumSMs * 6);
    k_classify_first<<<grid, kTileThreads, 0, L.stream>>>(rec, n, t0, piles, n_piles, P, p_cap, ev, ev_cap, hill_rec,
                                                          hill_cap, counters, status, ticket);
    L.count++;
}

void launch_fixpoint(Launch& L, Events ev, const uint32_t* n_events, uint32_t ev_cap, uint32_t* dbuf, uint32_t n_piles,
                     uint32_t* flags, uint32_t* counters, int coop_blocks) {
    void* args[] = {&ev, &n_events, &ev_cap, &dbuf, &n_piles, &flags, &counters};
    cudaLaunchCooperativeKernel((void*) k_containment_fixpoint, dim3(coop_blocks), dim3(256), args, 0, L.stream);
    L.count++;
}

int fixpoint_max_blocks() {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_containment_fixpoint, 256, 0);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int dev = 0, sms = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return per_sm * sms;
}

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
                           const uint2* piles, Events ev, uint32_t ev_cap, uint32_t* counters) {
    k_classify_final<<<grid_for(cap, 256, kNumSMs * 8), 256, 0, L.stream>>>(lst, n_ptr, cap, time_base, piles, ev, ev_cap,
                                                                            counters);
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
