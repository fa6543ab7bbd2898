// session.h — internal state of librala_b200.so shared by api.cu (single-GPU session, C ABI) and multi_api.cu
// (multi-rank orchestration over peer memory).  Not part of the public interface (include/rala_b200.h).
#pragma once

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rala_b200.h"
#include "kernels.h"
#include "lists.cuh"

using namespace rb;

struct rala_b200_ctx {
    bool owns_stream = true;
    cudaEvent_t ev[2]{};
    int device = 0;
    Launch L{nullptr, 0};
    std::string error;
    int coop_blocks = 0;
};

inline int fail(rala_b200_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    return code;
}

#define CU(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t err__ = (call);                                                                     \
        if (err__ != cudaSuccess)                                                                       \
            return fail((ctx), RALA_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), \
                        __FILE__, __LINE__);                                                            \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device buffer helper
// ---------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct ListBuf {
    DevBuf buf;
    List view{};
    cudaError_t reserve(uint32_t cap) {
        size_t col = align_up((size_t) cap * 4, 256);
        cudaError_t e = buf.reserve(col * 6 + align_up(cap, 256));
        if (e != cudaSuccess) return e;
        char* b = buf.as<char>();
        view.a = (uint32_t*) b;
        view.b = (uint32_t*) (b + col);
        view.ab = (uint32_t*) (b + 2 * col);
        view.ae = (uint32_t*) (b + 3 * col);
        view.bb = (uint32_t*) (b + 4 * col);
        view.be = (uint32_t*) (b + 5 * col);
        view.tag = (uint8_t*) (b + 6 * col);
        return cudaSuccess;
    }
};

enum Stage { ST_CLASSIFY = 0, ST_RETRIM, ST_FINALIZE, ST_BUILD, ST_TRANSITIVE, ST_K1_KERNEL, ST_K1B_KERNEL, ST_K3_KERNELS, ST_K1S_KERNEL };

struct rala_b200_graph {
    rala_b200_ctx* ctx = nullptr;
    // inputs
    DevBuf rec;        // the host's rala_ovl_t rows as uploaded (staging of the transpose)
    ListBuf recs;      // device-resident records: six columns, 24 B / record (classify.cu)
    DevBuf alive_bits; // one bit per pile: alive after the last containment resolution
    uint32_t n_rec = 0;
    DevBuf piles, piles_raw, pile_flags_raw, piles_initial;
    bool piles_fresh = false;       // set_piles since the last classify
    uint32_t n_piles = 0;
    DevBuf hills;   // 4 columns of n_hills: pile begin end cov
    uint32_t n_hills = 0;
    // lists
    uint32_t cap = 0;      // capacity of every overlap list
    uint32_t ev_cap = 0;   // capacity of the event arrays and victim segments (multi-GPU: events of ALL ranks)
    int rank = 0, world = 1;   // multi-GPU: which share of the source nodes the transitive phase takes
    uint32_t t0 = 0;       // global time (file position) of the first local record (multi-GPU shards)
    ListBuf ovl[2], inl[2];
    int ovl_cur = 0, inl_cur = 0;
    int slot_ovl = C_LIST0, slot_inl = C_LIST0 + 1, next_slot = C_LIST0 + 2;
    DevBuf events, hill_rec;
    DevBuf dbuf, flags, segs, tiles;   // dbuf: S | vcursor | vstart | work0 | work1 ; segs: seg_c | seg_t ; tiles: info | off
    DevBuf counters;
    DevBuf scan_pool;
    size_t scan_pool_words = 0, scan_used = 0;
    // graph
    DevBuf edges_aos;
    DevBuf seq_to_node, edges, row_ptr, cursor, col, col_eid, T, marked, heavy, work_counter;
    DevBuf edge_rank;   // row slot of every edge (k_emit_edges -> k_fill_csr)
    uint32_t edge_cap = 0, heavy_cap = 0, n_nodes_max = 0;
    // results written straight into the caller's memory by the run (rala_b200_graph_set_outputs)
    uint32_t* out_edges = nullptr;   // device-visible address of the caller's rala_edge_t rows
    uint8_t* out_marked = nullptr;
    uint32_t out_edges_cap = 0, out_marked_cap = 0;
    bool in_run = false;             // build is part of a whole run: the edge download may overlap the transitive pass
    bool download_pending = false;   // side stream 0 is still writing edges: joined at the end of the transitive stage
    // bookkeeping
    int final_time_base_slot = C_LIST0;   // counter slot holding the time base of `internals` in the final pass
    bool final_lists_ready = true;  // after finalize: have the filtered lists of graph.cpp:867-877 been written out?
    bool piles_dirty = true;        // pile table changed since the lists were last trimmed against it
    bool skip_clean_retrim = true;  // re-trimming against an unchanged table is the identity: skip the pass
    bool promote_pending = true;    // the table changed since the internals were last TYPED against it (retrim() trims but does not
                                    // re-type them): the next retrim_promote must run even if retrim() already cleared piles_dirty
    bool scan_pool_exhausted = false;   // a look-back kernel asked for more status words than the pool holds: the run is void
    int state = 0;                  // 0 empty, 1 inputs set, 2 classified, 3 finalized, 4 built, 5 reduced
    uint32_t retrim_passes = 0;
    cudaEvent_t ev_start[RALA_B200_N_STAGES]{}, ev_stop[RALA_B200_N_STAGES]{};
    bool ev_valid[RALA_B200_N_STAGES]{};
    // rala_b200_graph_run as a replayed CUDA graph (see run_key / RunGraph below)
    bool use_cuda_graph = true, capturing = false;
    struct RunGraph* run_graphs = nullptr;   // two cached instances: first run after set_piles / repeated run

    uint32_t* cnt() const { return counters.as<uint32_t>(); }
    Events events_view() const {
        Events e;
        size_t col = align_up((size_t) ev_cap * 4, 256);
        char* b = events.as<char>();
        e.v = (uint32_t*) b;
        e.c = (uint32_t*) (b + col);
        e.t = (uint32_t*) (b + 2 * col);
        return e;
    }
    GraphArrays graph_view() const {
        GraphArrays g;
        size_t col_b = align_up((size_t) edge_cap * 4, 256);
        g.seq_to_node = seq_to_node.as<uint32_t>();
        g.src = (uint32_t*) edges.as<char>();
        g.dst = (uint32_t*) (edges.as<char>() + col_b);
        g.len = (uint32_t*) (edges.as<char>() + 2 * col_b);
        g.row_ptr = row_ptr.as<uint32_t>();
        g.cursor = cursor.as<uint32_t>();
        g.col = col.as<uint2>();
        g.col_eid = col_eid.as<uint32_t>();
        g.rank = edge_rank.as<uint32_t>();
        g.T = T.as<uint8_t>();
        g.marked = marked.as<uint8_t>();
        return g;
    }
    HeavyItems heavy_view() const {
        HeavyItems h;
        size_t colb = align_up((size_t) heavy_cap * 4, 256);
        h.node = (uint32_t*) heavy.as<char>();
        h.hash_chunk = (uint32_t*) (heavy.as<char>() + colb);
        h.nbr_chunk = (uint32_t*) (heavy.as<char>() + 2 * colb);
        h.cap = heavy_cap;
        return h;
    }
    int new_slot() {
        int s = next_slot;
        next_slot = next_slot + 1 >= C_COUNT ? C_LIST0 : next_slot + 1;
        if (s == slot_ovl || s == slot_inl) return new_slot();
        return s;
    }
};


// ---- internals of api.cu used by the multi-rank orchestration (multi_api.cu) --------------------------------
size_t tiles_of(uint64_t n);
bool scan_state(rala_b200_graph* g, uint64_t n_max, unsigned long long** status, uint32_t** ticket);
bool stream_capturing(const rala_b200_graph* g);
cudaError_t stage_event(rala_b200_graph* g, cudaEvent_t ev);
cudaError_t begin_stage(rala_b200_graph* g, int stage);
cudaError_t end_stage(rala_b200_graph* g, int stage);
cudaError_t zero_counter(rala_b200_graph* g, int slot, int n = 1);
ResolveBufs resolve_bufs(const rala_b200_graph* g);
cudaError_t clear_victim_histogram(rala_b200_graph* g);
int reserve_events(rala_b200_graph* g, uint32_t ev_cap);
int reserve_edges(rala_b200_graph* g, uint32_t edge_cap);
int phase_events(rala_b200_graph* g);
int phase_survivors(rala_b200_graph* g);
int phase_final_events(rala_b200_graph* g, const uint32_t* ovl_base, const uint32_t* inl_base);
int read_counters(rala_b200_graph* g, uint32_t* h);
int materialize_final_lists(rala_b200_graph* g);
