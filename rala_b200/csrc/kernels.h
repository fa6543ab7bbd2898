// kernels.h — host-callable launchers of the hot-path kernels (internal to librala_b200.so).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lists.cuh"

namespace rb {


// slots of the device-side counter block (uint32 each)
enum Counter {
    C_P = 0,        // (unused)
    C_EV,           // containment events of the current resolution
    C_HILL,         // records touching a pile with hills
    C_ROUNDS,       // fixed-point rounds
    C_DSEL,         // which of the four death-time buffers holds the result
    C_NODES,        // nodes_.size()
    C_ALIVE,        // alive piles
    C_DOVETAILS,    // dovetail overlaps = edges / 2
    C_EDGES,        // edges_.size()
    C_PAIRS,        // marked pairs
    C_HEAVY,        // heavy work items of the transitive pass
    C_OVERFLOW,     // some list hit its capacity
    C_HOP_LO, C_HOP_HI,   // two-hop visits (64-bit)
    C_TBASE_OVL, C_TBASE_INL,     // multi-GPU: global time bases of the local lists in the final pass
    C_EV_FIRST, C_ROUNDS_FIRST,   // events / rounds of the first-pass resolution (C_EV / C_ROUNDS are reused by the final pass)
    C_LIST0,        // 16 rotating list-count slots follow
    C_COUNT = C_LIST0 + 16
};

struct Launch {
    cudaStream_t stream;
    uint64_t count;   // kernels launched so far
    // forked work inside one step (also inside a stream capture: the side streams join the capture through the
    // fork event and are joined back before it ends): [0] results written to the caller's memory while the
    // transitive pass runs, [1] kernels that do not depend on each other
    cudaStream_t side[2] = {nullptr, nullptr};
    cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
};

// side[k] continues from the current end of the main stream; false when the context has no side streams
inline bool fork_side(Launch& L, int k) {
    if (!L.side[k]) return false;
    if (cudaEventRecord(L.ev_fork[k], L.stream) != cudaSuccess) return false;
    return cudaStreamWaitEvent(L.side[k], L.ev_fork[k], 0) == cudaSuccess;
}
// the main stream waits for everything enqueued on side[k]
inline void join_side(Launch& L, int k) {
    cudaEventRecord(L.ev_join[k], L.side[k]);
    cudaStreamWaitEvent(L.stream, L.ev_join[k], 0);
}

// classify.cu
void launch_records_to_soa(Launch& L, const uint32_t* aos, uint32_t n, List recs);
void launch_unpack_records(Launch& L, const uint32_t* query_id, const uint32_t* group_end, uint32_t n_groups, const uint32_t* a_span,
                           const uint32_t* b_span, uint32_t n, List recs);
void launch_classify_events(Launch& L, List recs, uint32_t n, uint32_t t0, const uint2* piles, uint32_t n_piles,
                            Events ev, uint32_t ev_cap, uint32_t* vcount, uint32_t* hill_rec, uint32_t hill_cap, uint32_t* counters);
// runs of the survivors pass (128 records each): packed counts (overlaps | internals << 16) and the scanned offsets
struct RunBufs {
    uint32_t *cnt, *off_a, *off_b;   // num_runs each
};
void launch_classify_survivors(Launch& L, List recs, uint32_t n, const uint2* piles, const uint32_t* alive_bits, uint32_t n_piles,
                               List tmp_ovl, List tmp_inl, List ovl, uint32_t* n_ovl, List inl, uint32_t* n_inl, uint32_t cap,
                               RunBufs runs, unsigned long long* status, uint32_t* ticket);
uint32_t classify_num_runs(uint32_t n);
void launch_hill_coverage(Launch& L, List recs, uint32_t t0, const uint2* piles, const uint32_t* hill_rec,
                          uint32_t hill_cap, const uint32_t* hill_pile, const uint32_t* hill_begin,
                          const uint32_t* hill_end, uint32_t n_hills, uint32_t* hill_cov, const uint32_t* dbuf,
                          uint32_t n_piles, const uint32_t* counters);
// skip: optional device flag, non-zero = do nothing (multi-GPU: a containment pass without events); run_if: optional device
// count, zero = do nothing (single GPU: the final containment pass when it has no events)
void launch_apply_deaths(Launch& L, uint2* piles, uint32_t* dbuf, uint32_t n_piles, const uint32_t* counters,
                         uint32_t* alive_bits, bool decode, const uint32_t* skip = nullptr, const uint32_t* run_if = nullptr);
void launch_list_pass(Launch& L, int mode, List in, const uint32_t* n_in, uint32_t in_cap, const uint2* piles, List out_a,
                      uint32_t* n_out_a, List out_b, uint32_t* n_out_b, const uint32_t* b_base, uint32_t cap,
                      const uint32_t* dbuf, uint32_t n_piles, const uint32_t* time_base, const uint32_t* counters,
                      unsigned long long* status, uint32_t* ticket);
void launch_classify_final(Launch& L, List ovl, const uint32_t* n_ovl, const uint32_t* ovl_time_base, List inl,
                           const uint32_t* n_inl, const uint32_t* inl_time_base, uint32_t cap, const uint2* piles, Events ev,
                           uint32_t ev_cap, uint32_t* vcount, uint32_t* counters);
void launch_trim_classify_aos(Launch& L, uint32_t* rec, uint32_t n, const uint2* piles, uint32_t n_piles, uint8_t* type_out);
void launch_fill_u32(Launch& L, uint32_t* p, uint32_t v, size_t n);
void launch_pack_piles(Launch& L, const uint2* in, const uint8_t* flags, uint2* out, uint32_t n);
void launch_unpack_piles(Launch& L, const uint2* in, uint2* out, uint32_t n);
void launch_list_to_aos(Launch& L, List l, const uint32_t* n_ptr, uint32_t cap, uint32_t* out);
void launch_aos_to_list(Launch& L, const uint32_t* aos, uint32_t n, List l);
void launch_list_connections(Launch& L, List l, const uint32_t* n_ptr, uint32_t cap, uint32_t* out);

// containment.cu
struct ResolveBufs {
    uint32_t* S;         // n_piles: state during the resolution, death times (kInf = never) afterwards
    uint32_t* vcursor;   // n_piles + 1: per-victim event histogram on entry, scatter cursor inside
    uint32_t* vstart;    // n_piles + 1
    uint32_t *work0, *work1;   // n_piles each
    uint32_t* n_work;    // 4 words
    uint32_t *seg_c, *seg_t;   // event capacity each
};
int resolve_max_blocks();
// per-victim histogram of an imported event list (the classify kernels build it on the fly otherwise)
void launch_events_hist(Launch& L, Events ev, const uint32_t* n_events, uint32_t ev_cap, uint32_t* vcount);
// decode = false leaves the resolution's state encoding in rb.S: the following launch_apply_deaths(decode = true) turns
// it into death times while it applies them (one launch less when nothing reads the death times in between)
// skip_if_empty: with no events at all every kernel returns at once (nobody dies; rb.S is left as it is)
void launch_resolve(Launch& L, Events ev, const uint32_t* n_events, uint32_t ev_cap, ResolveBufs rb, uint32_t n_piles,
                    uint32_t* counters, unsigned long long* status, uint32_t* ticket, int coop_blocks, bool decode,
                    bool skip_if_empty = false);

// graph_build.cu
void launch_scan_u32(Launch& L, uint32_t* values_inout, uint32_t* exclusive_out, uint32_t n, unsigned long long* status,
                     uint32_t* ticket, const uint32_t* skip = nullptr, const uint32_t* run_if = nullptr);
struct GraphArrays {
    uint32_t* seq_to_node;   // n_piles
    uint32_t *src, *dst, *len;   // edge id order
    uint32_t* row_ptr;       // n_nodes_max + 1
    uint32_t* cursor;        // n_nodes_max + 1 (degree histogram, then fill cursor)
    uint2* col;              // CSR (dst, len)
    uint32_t* col_eid;       // CSR edge id
    uint32_t* rank;          // per edge: its position inside its source row, as handed out by the degree count in k_emit_edges
    uint8_t* T;              // per edge: transitive test passed
    uint8_t* marked;         // per edge: removed
};
void launch_node_ids(Launch& L, const uint2* piles, uint32_t n_piles, uint32_t* seq_to_node, uint32_t* counters,
                     unsigned long long* status, uint32_t* ticket);
// entries per thread of k_emit_edges (its tiles are 256 x this); its look-back scan needs one status word per tile.
// 2 and 1 (more, smaller tiles: a longer look-back chain) measured no faster than 4 (profiles/r02aa_ab.json).
#ifndef RB_EMIT_ITEMS
#define RB_EMIT_ITEMS 4
#endif
constexpr int kEmitItems = RB_EMIT_ITEMS;
inline uint64_t emit_scan_span(uint64_t cap) { return cap * (4 / kEmitItems); }   // in units of the common 1024-entry tile
void launch_emit_edges(Launch& L, List ovl, const uint32_t* n_ptr, uint32_t cap, const uint2* piles, GraphArrays g,
                       uint32_t edge_cap, uint32_t* counters, unsigned long long* status, uint32_t* ticket);
void launch_pack_edges(Launch& L, cudaStream_t stream, GraphArrays g, uint32_t edge_cap, const uint32_t* n_edges_ptr, uint32_t* out);
// degree histogram for an edge list that did not come from emit_edges (stateless transitive stage)
void launch_degree_hist(Launch& L, const uint32_t* src, const uint32_t* n_edges_ptr, uint32_t edge_cap, uint32_t* cursor);
void launch_export_padded(Launch& L, const uint32_t* c0, const uint32_t* c1, const uint32_t* c2, const uint32_t* n_ptr,
                          uint32_t src_cap, uint32_t cap, uint32_t* block);
void launch_import_gathered(Launch& L, const uint32_t* gathered, uint32_t cap, uint32_t world, uint32_t* d0, uint32_t* d1,
                            uint32_t* d2, uint32_t dst_cap, uint32_t* n_out, uint32_t* overflow);
void launch_time_bases(Launch& L, const uint32_t* counts, uint32_t rank, uint32_t world, uint32_t* bases);
// host-facing adjacency view: ids of the (unmarked) edges per node, ascending edge id; key = src (suffix) or dst (prefix) column
void launch_adjacency_view(Launch& L, const uint32_t* key, const uint8_t* marked, const uint32_t* n_edges_ptr, uint32_t edge_cap,
                           uint32_t n_nodes, uint32_t* degree_cursor, uint32_t* row_ptr, uint32_t* ids_tmp, uint32_t* ids_sorted,
                           unsigned long long* status, uint32_t* ticket);
void launch_build_csr(Launch& L, GraphArrays g, uint32_t n_nodes_max, uint32_t edge_cap, uint32_t* counters,
                      unsigned long long* status, uint32_t* ticket, bool ranked = false);

// transitive.cu
struct HeavyItems {
    uint32_t *node, *hash_chunk, *nbr_chunk;
    uint32_t cap;
};
void launch_transitive(Launch& L, GraphArrays g, uint32_t n_nodes_max, uint32_t edge_cap, HeavyItems heavy,
                       uint32_t* work_counter, uint32_t* counters, uint32_t node_begin, uint32_t node_end,
                       const uint32_t* node_range /* device {begin, end}, nullable */);
void launch_node_range(Launch& L, GraphArrays g, const uint32_t* counters, uint32_t rank, uint32_t world, uint32_t* out);
void launch_finalize_marks(Launch& L, GraphArrays g, uint32_t edge_cap, uint32_t* counters, uint8_t* marked_copy /* nullable */,
                           uint32_t copy_cap);

// force-load every kernel of a translation unit (CUDA's lazy loading otherwise loads at the first launch, which waits
// for the device to drain: a deadlock next to a spinning barrier kernel, see fabric.cu)
void preload_classify();
void preload_containment();
void preload_graph_build();
void preload_transitive();
void preload_fabric();
void preload_frontend();

// frontend.cu
void launch_filter_duplicates(Launch& L, const uint32_t* a, const uint32_t* b, const uint32_t* len, uint32_t n, uint8_t* valid);

}  // namespace rb
