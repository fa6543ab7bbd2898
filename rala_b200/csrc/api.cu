// api.cu — the C ABI (include/rala_b200.h): context, stateless stages, and the graph session that
// chains the kernels of classify.cu / graph_build.cu / transitive.cu with no host synchronisation
// between stages (every data-dependent size stays on the device).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rala_b200.h"
#include "session.h"

using namespace rb;

extern "C" int rala_b200_abi_version(void) { return RALA_B200_ABI_VERSION; }

extern "C" int rala_b200_create(rala_b200_ctx** out, int device) {
    if (!out) return RALA_B200_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) return RALA_B200_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RALA_B200_ERR_NO_DEVICE;
    if (prop.major != 10) return RALA_B200_ERR_NO_DEVICE;   // sm_100a SASS only: no other target, no fallback
    rala_b200_ctx* ctx = new rala_b200_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->L.stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return RALA_B200_ERR_CUDA;
    }
    ctx->coop_blocks = resolve_max_blocks();
    preload_classify();
    preload_containment();
    preload_graph_build();
    preload_transitive();
    preload_fabric();
    preload_frontend();
    cudaGetLastError();
    cudaEventCreate(&ctx->ev[0]);
    cudaEventCreate(&ctx->ev[1]);
    for (int k = 0; k < 2; ++k) {   // forked work inside a step (kernels.h Launch); without them everything stays on the main stream
        if (cudaStreamCreateWithFlags(&ctx->L.side[k], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->L.ev_fork[k], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->L.ev_join[k], cudaEventDisableTiming) != cudaSuccess) {
            ctx->L.side[k] = nullptr;
            cudaGetLastError();
        }
    }
    *out = ctx;
    return RALA_B200_OK;
}

extern "C" int rala_b200_create_on_stream(rala_b200_ctx** out, int device, void* cuda_stream) {
    int rc = rala_b200_create(out, device);
    if (rc) return rc;
    cudaStreamDestroy((*out)->L.stream);
    (*out)->L.stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    (*out)->owns_stream = false;
    return RALA_B200_OK;
}

extern "C" void rala_b200_destroy(rala_b200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->L.stream && ctx->owns_stream) cudaStreamDestroy(ctx->L.stream);
    for (int k = 0; k < 2; ++k) {
        if (ctx->L.side[k]) cudaStreamDestroy(ctx->L.side[k]);
        if (ctx->L.ev_fork[k]) cudaEventDestroy(ctx->L.ev_fork[k]);
        if (ctx->L.ev_join[k]) cudaEventDestroy(ctx->L.ev_join[k]);
    }
    delete ctx;
}

extern "C" int rala_b200_event_record(rala_b200_ctx* ctx, int which) {
    if (!ctx || which < 0 || which > 1) return RALA_B200_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaEventRecord(ctx->ev[which], ctx->L.stream));
    return RALA_B200_OK;
}

extern "C" int rala_b200_event_elapsed_ms(rala_b200_ctx* ctx, float* ms) {
    if (!ctx || !ms) return RALA_B200_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaEventSynchronize(ctx->ev[1]));
    CU(ctx, cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
    return RALA_B200_OK;
}

extern "C" int rala_b200_synchronize(rala_b200_ctx* ctx) {
    if (!ctx) return RALA_B200_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    return RALA_B200_OK;
}

extern "C" const char* rala_b200_last_error(const rala_b200_ctx* ctx) { return ctx ? ctx->error.c_str() : "no context"; }
extern "C" uint64_t rala_b200_launch_count(const rala_b200_ctx* ctx) { return ctx ? ctx->L.count : 0; }

size_t tiles_of(uint64_t n) { return (size_t) ((n + kTile - 1) / kTile) + 1; }

// a fresh, zeroed (status words, ticket) pair from the pool for one look-back kernel
bool scan_state(rala_b200_graph* g, uint64_t n_max, unsigned long long** status, uint32_t** ticket) {
    size_t words = tiles_of(n_max) + 1;
    if (g->scan_used + words > g->scan_pool_words) {
        // Handing out words an earlier look-back kernel of the same stage already wrote would give wrong prefixes (or spin).
        // The pool is sized for every stage (reserve_scan_pool); should a new caller break that, the run is declared
        // void (read_counters reports it) and the words are recycled only so that the kernel arguments stay valid.
        g->scan_pool_exhausted = true;
        g->scan_used = 0;
    }
    unsigned long long* base = g->scan_pool.as<unsigned long long>() + g->scan_used;
    *ticket = reinterpret_cast<uint32_t*>(base);
    *status = base + 1;
    g->scan_used += words;
    return !g->scan_pool_exhausted;
}

// stage timers: plain event records, left out of a stream capture (an event recorded inside a capture cannot be timed)
// `capturing` covers the library's own capture (rala_b200_graph_run); a caller may also capture the phase calls
// into a graph of its own (rala_b200/multi.py: kernels + NCCL collectives of one multi-GPU step), so ask the stream.
bool stream_capturing(const rala_b200_graph* g) {
    if (g->capturing) return true;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(g->ctx->L.stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

cudaError_t stage_event(rala_b200_graph* g, cudaEvent_t ev) {
    return stream_capturing(g) ? cudaSuccess : cudaEventRecord(ev, g->ctx->L.stream);
}

cudaError_t begin_stage(rala_b200_graph* g, int stage) {
    g->scan_used = 0;
    cudaError_t e = cudaMemsetAsync(g->scan_pool.p, 0, g->scan_pool_words * 8, g->ctx->L.stream);
    if (e != cudaSuccess) return e;
    return stage_event(g, g->ev_start[stage]);
}

cudaError_t end_stage(rala_b200_graph* g, int stage) {
    g->ev_valid[stage] = !stream_capturing(g);
    return stage_event(g, g->ev_stop[stage]);
}

cudaError_t zero_counter(rala_b200_graph* g, int slot, int n) {
    return cudaMemsetAsync(g->cnt() + slot, 0, 4 * (size_t) n, g->ctx->L.stream);
}

extern "C" int rala_b200_graph_create(rala_b200_ctx* ctx, rala_b200_graph** out) {
    if (!ctx || !out) return RALA_B200_ERR_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    rala_b200_graph* g = new rala_b200_graph();
    g->ctx = ctx;
    for (int i = 0; i < RALA_B200_N_STAGES; ++i) {
        cudaEventCreate(&g->ev_start[i]);
        cudaEventCreate(&g->ev_stop[i]);
    }
    cudaError_t e = g->counters.reserve(C_COUNT * 4);
    if (e == cudaSuccess) e = g->flags.reserve(64);
    if (e == cudaSuccess) e = g->work_counter.reserve(64);
    if (e == cudaSuccess) e = cudaMemset(g->counters.p, 0, C_COUNT * 4);
    if (e != cudaSuccess) {
        delete g;
        return fail(ctx, RALA_B200_ERR_CUDA, "graph_create: %s", cudaGetErrorString(e));
    }
    *out = g;
    return RALA_B200_OK;
}

static void drop_run_graphs(rala_b200_graph* g);

extern "C" void rala_b200_graph_destroy(rala_b200_graph* g) {
    if (!g) return;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->L.stream);
    drop_run_graphs(g);
    DevBuf* bufs[] = {&g->rec, &g->recs.buf, &g->alive_bits, &g->piles, &g->piles_raw, &g->pile_flags_raw, &g->piles_initial, &g->hills, &g->ovl[0].buf,
                      &g->ovl[1].buf, &g->inl[0].buf, &g->inl[1].buf, &g->events, &g->hill_rec, &g->dbuf, &g->flags, &g->segs, &g->tiles,
                      &g->counters, &g->scan_pool, &g->seq_to_node, &g->edges, &g->row_ptr, &g->cursor, &g->col,
                      &g->col_eid, &g->T, &g->marked, &g->heavy, &g->work_counter, &g->edges_aos};
    for (DevBuf* b : bufs) b->release();
    for (int i = 0; i < RALA_B200_N_STAGES; ++i) {
        cudaEventDestroy(g->ev_start[i]);
        cudaEventDestroy(g->ev_stop[i]);
    }
    delete g;
}

int reserve_events(rala_b200_graph* g, uint32_t ev_cap) {
    if (ev_cap <= g->ev_cap) return RALA_B200_OK;
    rala_b200_ctx* ctx = g->ctx;
    // growing re-allocates: contents are not preserved (callers grow before they fill)
    g->ev_cap = ev_cap;
    CU(ctx, g->events.reserve(align_up((size_t) ev_cap * 4, 256) * 3));
    CU(ctx, g->segs.reserve(align_up((size_t) ev_cap * 4, 256) * 2));
    return RALA_B200_OK;
}

int reserve_edges(rala_b200_graph* g, uint32_t edge_cap) {
    if (edge_cap <= g->edge_cap) return RALA_B200_OK;
    rala_b200_ctx* ctx = g->ctx;
    g->edge_cap = edge_cap;
    CU(ctx, g->edges.reserve(align_up((size_t) edge_cap * 4, 256) * 3));
    CU(ctx, g->col.reserve((size_t) edge_cap * 8));
    CU(ctx, g->col_eid.reserve((size_t) edge_cap * 4));
#if RB_OPT_RANK
    CU(ctx, g->edge_rank.reserve((size_t) edge_cap * 4));
#endif
    CU(ctx, g->T.reserve(align_up(edge_cap, 256)));
    CU(ctx, g->marked.reserve(align_up(edge_cap, 256)));
    g->heavy_cap = edge_cap / 16 + 4096;
    CU(ctx, g->heavy.reserve(align_up((size_t) g->heavy_cap * 4, 256) * 3));
    return RALA_B200_OK;
}

static int reserve_scan_pool(rala_b200_graph* g) {
    // worst stage: classify = runs scan + two containment scans over the piles; build = piles + list + node degrees;
    // the multi-GPU session adds scans over its exchange capacities (<= 2 x the record count): 16 x leaves room
    size_t words = 16 * (tiles_of(g->n_rec) + tiles_of(2ull * g->n_piles + 8) + 8);
    if (words > g->scan_pool_words) {
        CU(g->ctx, g->scan_pool.reserve(words * 8));
        g->scan_pool_words = words;
    }
    return RALA_B200_OK;
}

// device lists sized for n records (worst case every record survives; the scratch lists of the survivors pass are
// indexed by record position)
static int reserve_for_records(rala_b200_graph* g, uint64_t n) {
    rala_b200_ctx* ctx = g->ctx;
    g->n_rec = (uint32_t) n;
    CU(ctx, g->recs.reserve((uint32_t) ((n + 3) / 4 * 4 + 4)));
    uint32_t cap = (uint32_t) align_up((size_t) n, 128);
    if (cap < 1024) cap = 1024;
    if (cap > g->cap) {
        for (int i = 0; i < 2; ++i) {
            CU(ctx, g->ovl[i].reserve(cap));
            CU(ctx, g->inl[i].reserve(cap));
        }
        CU(ctx, g->hill_rec.reserve((size_t) cap * 4));
        g->cap = cap;
        int rc2 = reserve_events(g, cap);
        if (!rc2) rc2 = reserve_edges(g, 2 * cap);
        if (rc2) return rc2;
    }
    CU(ctx, g->tiles.reserve(align_up((size_t) classify_num_runs((uint32_t) n) + 8, 64) * 4 * 3));
    int rc = reserve_scan_pool(g);
    if (rc) return rc;
    if (g->state < 1 && g->n_piles) g->state = 1;
    if (g->state > 1) g->state = 1;
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_set_overlaps(rala_b200_graph* g, const rala_ovl_t* ovl, uint64_t n) {
    if (!g || (n && !ovl)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (n >= (1ull << 31)) return fail(ctx, RALA_B200_ERR_LIMIT, "too many overlap records (%llu >= 2^31)", (unsigned long long) n);
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, g->rec.reserve(align_up((size_t) n * sizeof(rala_ovl_t) + 16, 256)));
    if (n) CU(ctx, cudaMemcpyAsync(g->rec.p, ovl, (size_t) n * sizeof(rala_ovl_t), cudaMemcpyHostToDevice, ctx->L.stream));
    int rc = reserve_for_records(g, n);
    if (rc) return rc;
    // layout in HBM: rows -> six columns, once per upload (the kernels read 16 bytes per column and thread)
    launch_records_to_soa(ctx->L, g->rec.as<uint32_t>(), (uint32_t) n, g->recs.view);
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

// The same records as six host columns in the device layout: 24 bytes per record cross PCIe instead of 28 and
// land where the kernels read them (no staging buffer, no transpose kernel).
extern "C" int rala_b200_graph_set_overlaps_columns(rala_b200_graph* g, const uint32_t* a_id, const uint32_t* b_id,
                                                    const uint32_t* a_begin, const uint32_t* a_end, const uint32_t* b_begin,
                                                    const uint32_t* b_end, uint64_t n) {
    if (!g || (n && (!a_id || !b_id || !a_begin || !a_end || !b_begin || !b_end))) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (n >= (1ull << 31)) return fail(ctx, RALA_B200_ERR_LIMIT, "too many overlap records (%llu >= 2^31)", (unsigned long long) n);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = reserve_for_records(g, n);
    if (rc) return rc;
    if (n) {
        const List& r = g->recs.view;
        const uint32_t* src[6] = {a_id, b_id, a_begin, a_end, b_begin, b_end};
        uint32_t* dst[6] = {r.a, r.b, r.ab, r.ae, r.bb, r.be};
        for (int k = 0; k < 6; ++k)
            CU(ctx, cudaMemcpyAsync(dst[k], src[k], (size_t) n * 4, cudaMemcpyHostToDevice, ctx->L.stream));
    }
    return RALA_B200_OK;
}

// Compact upload (12 B / record): b_id lands in place, the two span columns and the group table go through the staging
// buffer and are expanded by one kernel.
extern "C" int rala_b200_graph_set_overlaps_packed(rala_b200_graph* g, const uint32_t* query_id, const uint32_t* group_end, uint32_t n_groups,
                                                   const uint32_t* b_id, const uint32_t* a_span, const uint32_t* b_span, uint64_t n) {
    if (!g || (n && (!query_id || !group_end || !b_id || !a_span || !b_span || !n_groups))) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (n >= (1ull << 31)) return fail(ctx, RALA_B200_ERR_LIMIT, "too many overlap records (%llu >= 2^31)", (unsigned long long) n);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = reserve_for_records(g, n);
    if (rc) return rc;
    if (!n) return RALA_B200_OK;
    const size_t col = align_up(((size_t) n + 4) * 4, 256), grp = align_up((size_t) n_groups * 4, 256);
    CU(ctx, g->rec.reserve(2 * col + 2 * grp));
    char* st = g->rec.as<char>();
    uint32_t *d_a = (uint32_t*) st, *d_b = (uint32_t*) (st + col), *d_q = (uint32_t*) (st + 2 * col), *d_e = (uint32_t*) (st + 2 * col + grp);
    CU(ctx, cudaMemcpyAsync(g->recs.view.b, b_id, (size_t) n * 4, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx, cudaMemcpyAsync(d_a, a_span, (size_t) n * 4, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx, cudaMemcpyAsync(d_b, b_span, (size_t) n * 4, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx, cudaMemcpyAsync(d_q, query_id, (size_t) n_groups * 4, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx, cudaMemcpyAsync(d_e, group_end, (size_t) n_groups * 4, cudaMemcpyHostToDevice, ctx->L.stream));
    launch_unpack_records(ctx->L, d_q, d_e, n_groups, d_a, d_b, (uint32_t) n, g->recs.view);
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

// Results straight into the caller's memory (pinned host memory or device memory the GPU can address).
extern "C" int rala_b200_graph_set_outputs(rala_b200_graph* g, rala_edge_t* edges_out, uint64_t edges_cap, uint8_t* marked_out,
                                           uint64_t marked_cap) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    void* dev[2] = {nullptr, nullptr};
    const void* host[2] = {edges_out, marked_out};
    for (int k = 0; k < 2; ++k) {
        if (!host[k]) continue;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, host[k]) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered || !attr.devicePointer) {
            cudaGetLastError();
            return fail(ctx, RALA_B200_ERR_ARG, "set_outputs: the %s buffer is not addressable by the GPU (use cudaHostAlloc / cudaHostRegister "
                        "memory, or get_edges / get_marked for pageable memory)", k ? "marks" : "edge");
        }
        dev[k] = attr.devicePointer;
    }
    if (edges_cap >= (1ull << 31)) edges_cap = (1ull << 31) - 1;
    if (marked_cap >= (1ull << 31)) marked_cap = (1ull << 31) - 1;
    g->out_edges = reinterpret_cast<uint32_t*>(dev[0]);
    g->out_edges_cap = edges_out ? (uint32_t) edges_cap : 0u;
    g->out_marked = reinterpret_cast<uint8_t*>(dev[1]);
    g->out_marked_cap = marked_out ? (uint32_t) marked_cap : 0u;
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_set_piles(rala_b200_graph* g, const rala_pile_t* piles, const uint8_t* flags,
                                         uint32_t n_piles) {
    if (!g || (n_piles && !piles)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (n_piles >= (1u << 31)) return fail(ctx, RALA_B200_ERR_LIMIT, "too many piles");
    // event_code (common.cuh) and the packed table (end | flags << 30) rely on the documented limit: valid regions end below 2^30
    // (two branch-free passes over the table as 64-bit words, begin in the low half: it is re-uploaded for every batch; the
    // offender is only looked for when there is one)
    {
        static_assert(sizeof(rala_pile_t) == 8, "a pile is one 64-bit word: begin | end << 32");
        unsigned long long any = 0, inverted = 0;
        for (uint32_t i = 0; i < n_piles; ++i) {
            unsigned long long w;
            memcpy(&w, &piles[i], 8);
            any |= w;
            inverted += (unsigned long long) (((uint32_t) w > (uint32_t) (w >> 32)) & ((w >> 32) != 0));
        }
        if ((any >> 62) || inverted) {
            for (uint32_t i = 0; i < n_piles; ++i) {
                if (piles[i].end >= (1u << 30)) return fail(ctx, RALA_B200_ERR_LIMIT, "pile %u ends at %u: read lengths must be < 2^30", i, piles[i].end);
                if (piles[i].end && piles[i].begin > piles[i].end) return fail(ctx, RALA_B200_ERR_ARG, "pile %u: begin %u > end %u", i, piles[i].begin, piles[i].end);
            }
        }
    }
    CU(ctx, g->piles.reserve((size_t) n_piles * 8 + 16));
    CU(ctx, g->piles_raw.reserve((size_t) n_piles * 8 + 16));
    CU(ctx, g->piles_initial.reserve((size_t) n_piles * 8 + 16));
    CU(ctx, g->pile_flags_raw.reserve((size_t) n_piles + 16));
    CU(ctx, cudaMemcpyAsync(g->piles_raw.p, piles, (size_t) n_piles * 8, cudaMemcpyHostToDevice, ctx->L.stream));
    if (flags) CU(ctx, cudaMemcpyAsync(g->pile_flags_raw.p, flags, n_piles, cudaMemcpyHostToDevice, ctx->L.stream));
    launch_pack_piles(ctx->L, g->piles_raw.as<uint2>(), flags ? g->pile_flags_raw.as<uint8_t>() : nullptr,
                      g->piles.as<uint2>(), n_piles);
    if (n_piles != g->n_piles) {
        g->n_piles = n_piles;
        g->n_nodes_max = 2 * n_piles;
        CU(ctx, g->dbuf.reserve(align_up((size_t) n_piles + 64, 64) * 4 * 5));
        CU(ctx, g->alive_bits.reserve(((size_t) n_piles / 32 + 2) * 4));
        CU(ctx, g->seq_to_node.reserve((size_t) n_piles * 4 + 16));
        CU(ctx, g->row_ptr.reserve(((size_t) g->n_nodes_max + 8) * 4));
        CU(ctx, g->cursor.reserve(((size_t) g->n_nodes_max + 8) * 4));
        int rc = reserve_scan_pool(g);
        if (rc) return rc;
    }
    g->piles_dirty = true;
    g->promote_pending = true;
    g->piles_fresh = true;
    if (g->state < 1 && g->n_rec) g->state = 1;
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_set_hills(rala_b200_graph* g, const rala_hill_t* hills, uint32_t n_hills) {
    if (!g || (n_hills && !hills)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    g->n_hills = n_hills;
    if (!n_hills) return RALA_B200_OK;
    std::vector<uint32_t> cols((size_t) n_hills * 4, 0u);
    for (uint32_t i = 0; i < n_hills; ++i) {
        if (i && hills[i].pile < hills[i - 1].pile) return fail(ctx, RALA_B200_ERR_ARG, "hills must be grouped by ascending pile id");
        cols[i] = hills[i].pile;
        cols[n_hills + i] = hills[i].begin;
        cols[2 * (size_t) n_hills + i] = hills[i].end;
    }
    CU(ctx, g->hills.reserve(cols.size() * 4));
    CU(ctx, cudaMemcpyAsync(g->hills.p, cols.data(), cols.size() * 4, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));   // cols is a stack-owned staging vector
    return RALA_B200_OK;
}

ResolveBufs resolve_bufs(const rala_b200_graph* g) {
    ResolveBufs r;
    const size_t stride = align_up((size_t) g->n_piles + 64, 64);   // 256-byte aligned sub-arrays (16-byte vector loads)
    uint32_t* b = g->dbuf.as<uint32_t>();
    r.S = b;
    r.vcursor = b + stride;
    r.vstart = b + 2 * stride;
    r.work0 = b + 3 * stride;
    r.work1 = b + 4 * stride;
    r.n_work = g->flags.as<uint32_t>();
    r.seg_c = g->segs.as<uint32_t>();
    r.seg_t = (uint32_t*) (g->segs.as<char>() + align_up((size_t) g->ev_cap * 4, 256));
    return r;
}

// the classify kernels bump the per-victim histogram while they emit events: clear it first
cudaError_t clear_victim_histogram(rala_b200_graph* g) {
    return cudaMemsetAsync(resolve_bufs(g).vcursor, 0, ((size_t) g->n_piles + 64) * 4, g->ctx->L.stream);
}

static int resolve_containment(rala_b200_graph* g, bool decode, bool skip_if_empty) {
    rala_b200_ctx* ctx = g->ctx;
    unsigned long long* status;
    uint32_t* ticket;
    scan_state(g, (uint64_t) g->n_piles + 1, &status, &ticket);
    launch_resolve(ctx->L, g->events_view(), g->cnt() + C_EV, g->ev_cap, resolve_bufs(g), g->n_piles, g->cnt(), status, ticket,
                   ctx->coop_blocks, decode, skip_if_empty);
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

// ---- graph.cpp:443-518 in three phases (the multi-GPU path exchanges events between them) ----------------
int phase_events(rala_b200_graph* g) {
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 1) return fail(ctx, RALA_B200_ERR_STATE, "classify: set_overlaps and set_piles first");
    CU(ctx, cudaSetDevice(ctx->device));
    // classify kills piles in the working table: keep the table it started from, so that calling classify
    // again without a new set_piles re-runs on the same input (bench loops, retries)
    if (g->piles_fresh)
        CU(ctx, cudaMemcpyAsync(g->piles_initial.p, g->piles.p, (size_t) g->n_piles * 8, cudaMemcpyDeviceToDevice, ctx->L.stream));
    else
        CU(ctx, cudaMemcpyAsync(g->piles.p, g->piles_initial.p, (size_t) g->n_piles * 8, cudaMemcpyDeviceToDevice, ctx->L.stream));
    g->piles_fresh = false;
    CU(ctx, begin_stage(g, ST_CLASSIFY));
    CU(ctx, cudaMemsetAsync(g->counters.p, 0, C_COUNT * 4, ctx->L.stream));
    if (g->n_hills) CU(ctx, cudaMemsetAsync(g->hills.as<uint32_t>() + 3 * (size_t) g->n_hills, 0, (size_t) g->n_hills * 4, ctx->L.stream));
    CU(ctx, clear_victim_histogram(g));
    CU(ctx, stage_event(g, g->ev_start[ST_K1_KERNEL]));
    launch_classify_events(ctx->L, g->recs.view, g->n_rec, g->t0, g->piles.as<uint2>(), g->n_piles, g->events_view(),
                           g->ev_cap, resolve_bufs(g).vcursor, g->hill_rec.as<uint32_t>(), g->cap, g->cnt());
    CU(ctx, end_stage(g, ST_K1_KERNEL));
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

static int phase_resolve(rala_b200_graph* g, bool first_pass) {
    rala_b200_ctx* ctx = g->ctx;
    // the hill counters read the death times BEFORE the deaths are applied to the pile table; without hills nothing
    // sits between the resolution and k_apply_deaths, which then decodes the states itself
    const bool fused_decode = RB_OPT_FUSE && !(first_pass && g->n_hills);
    CU(ctx, stage_event(g, g->ev_start[ST_K1B_KERNEL]));
    // The final pass (graph.cpp:831-866) has no events at all on clean data: scan, scatter, resolution and the sweep over the
    // pile table then return at once (nobody dies; the death times of the first pass, all "never" for the piles still
    // in the lists, stay in place).  The first pass always runs: it also builds the liveness bitmap.
    const bool skip_if_empty = RB_OPT_FUSE && !first_pass && fused_decode;
    int rc = resolve_containment(g, !fused_decode, skip_if_empty);
    if (rc) return rc;
    CU(ctx, end_stage(g, ST_K1B_KERNEL));
    if (first_pass && g->n_hills) {
        const uint32_t* h = g->hills.as<uint32_t>();
        launch_hill_coverage(ctx->L, g->recs.view, g->t0, g->piles.as<uint2>(), g->hill_rec.as<uint32_t>(), g->cap, h,
                             h + g->n_hills, h + 2 * (size_t) g->n_hills, g->n_hills,
                             g->hills.as<uint32_t>() + 3 * (size_t) g->n_hills, g->dbuf.as<uint32_t>(), g->n_piles, g->cnt());
    }
    launch_apply_deaths(ctx->L, g->piles.as<uint2>(), g->dbuf.as<uint32_t>(), g->n_piles, g->cnt(), g->alive_bits.as<uint32_t>(), fused_decode,
                        nullptr, skip_if_empty ? g->cnt() + C_EV : nullptr);
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

int phase_survivors(rala_b200_graph* g) {
    rala_b200_ctx* ctx = g->ctx;
    g->ovl_cur = 0;
    g->inl_cur = 0;
    g->slot_ovl = C_LIST0;
    g->slot_inl = C_LIST0 + 1;
    g->next_slot = C_LIST0 + 2;
    CU(ctx, zero_counter(g, C_LIST0, 2));
    CU(ctx, stage_event(g, g->ev_start[ST_K1S_KERNEL]));
    {
        const size_t nr = align_up((size_t) classify_num_runs(g->n_rec) + 8, 64);
        uint32_t* t = g->tiles.as<uint32_t>();
        RunBufs runs{t, t + nr, t + 2 * nr};
        unsigned long long* status;
        uint32_t* ticket;
        scan_state(g, nr, &status, &ticket);
        launch_classify_survivors(ctx->L, g->recs.view, g->n_rec, g->piles.as<uint2>(), g->alive_bits.as<uint32_t>(), g->n_piles,
                                  g->ovl[1].view, g->inl[1].view, g->ovl[0].view, g->cnt() + g->slot_ovl, g->inl[0].view,
                                  g->cnt() + g->slot_inl, g->cap, runs, status, ticket);
    }
    CU(ctx, end_stage(g, ST_K1S_KERNEL));
    CU(ctx, cudaGetLastError());
    CU(ctx, end_stage(g, ST_CLASSIFY));
    g->piles_dirty = false;   // lists are trimmed against the table as it stands (only liveness changed, and the split filtered on it)
    g->promote_pending = false;   // and every internal was typed against it
    g->state = 2;
    g->final_lists_ready = true;
    g->retrim_passes = 0;
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_classify(rala_b200_graph* g) {
    if (!g) return RALA_B200_ERR_ARG;
    int rc = phase_events(g);
    if (!rc) rc = phase_resolve(g, true);
    if (!rc) rc = phase_survivors(g);
    return rc;
}

static int retrim_list(rala_b200_graph* g, ListBuf* bufs, int* cur, int* slot) {
    rala_b200_ctx* ctx = g->ctx;
    unsigned long long* status;
    uint32_t* ticket;
    scan_state(g, g->cap, &status, &ticket);
    int out_slot = g->new_slot();
    CU(ctx, zero_counter(g, out_slot));
    List none{};
    launch_list_pass(ctx->L, 1 /*kRetrim*/, bufs[*cur].view, g->cnt() + *slot, g->cap, g->piles.as<uint2>(), bufs[*cur ^ 1].view,
                     g->cnt() + out_slot, none, nullptr, nullptr, g->cap, nullptr, g->n_piles, nullptr, g->cnt(), status, ticket);
    *cur ^= 1;
    *slot = out_slot;
    return RALA_B200_OK;
}

// graph.cpp:722-736
extern "C" int rala_b200_graph_retrim(rala_b200_graph* g) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state != 2) return fail(ctx, RALA_B200_ERR_STATE, "retrim: classify first");
    if (!g->piles_dirty && g->skip_clean_retrim) return RALA_B200_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, begin_stage(g, ST_RETRIM));
    int rc = retrim_list(g, g->ovl, &g->ovl_cur, &g->slot_ovl);
    if (!rc) rc = retrim_list(g, g->inl, &g->inl_cur, &g->slot_inl);
    if (rc) return rc;
    CU(ctx, cudaGetLastError());
    CU(ctx, end_stage(g, ST_RETRIM));
    g->piles_dirty = false;
    g->retrim_passes += 1;
    return RALA_B200_OK;
}

// graph.cpp:801-824
extern "C" int rala_b200_graph_retrim_promote(rala_b200_graph* g, int* is_changed) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state != 2) return fail(ctx, RALA_B200_ERR_STATE, "retrim_promote: classify first");
    if (is_changed) *is_changed = 0;
    // Against a pile table the internals were already trimmed AND typed with, trim() is the identity and they keep their
    // (non-dovetail) type.  retrim() trims them but does not re-type them: after set_piles -> retrim the promotion of
    // graph.cpp:809-823 still has to run (promote_pending), although piles_dirty is clear.
    if (!g->piles_dirty && !g->promote_pending && g->skip_clean_retrim) return RALA_B200_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, begin_stage(g, ST_RETRIM));
    int before_slot = g->slot_ovl;
    int rc = retrim_list(g, g->ovl, &g->ovl_cur, &g->slot_ovl);
    if (rc) return rc;
    int after_slot = g->slot_ovl;
    // internals: survivors stay (stream A), new dovetails are appended to `overlaps` (stream B)
    unsigned long long* status;
    uint32_t* ticket;
    scan_state(g, g->cap, &status, &ticket);
    int inl_out = g->new_slot(), ovl_out = g->new_slot();
    CU(ctx, zero_counter(g, inl_out));
    CU(ctx, cudaMemcpyAsync(g->cnt() + ovl_out, g->cnt() + after_slot, 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    launch_list_pass(ctx->L, 2 /*kPromote*/, g->inl[g->inl_cur].view, g->cnt() + g->slot_inl, g->cap, g->piles.as<uint2>(),
                     g->inl[g->inl_cur ^ 1].view, g->cnt() + inl_out, g->ovl[g->ovl_cur].view, g->cnt() + ovl_out,
                     g->cnt() + after_slot, g->cap, nullptr, g->n_piles, nullptr, g->cnt(), status, ticket);
    g->inl_cur ^= 1;
    g->slot_inl = inl_out;
    g->slot_ovl = ovl_out;
    CU(ctx, cudaGetLastError());
    CU(ctx, end_stage(g, ST_RETRIM));
    g->piles_dirty = false;
    g->promote_pending = false;
    g->retrim_passes += 1;
    if (is_changed) {
        uint32_t h[C_COUNT];
        CU(ctx, cudaMemcpyAsync(h, g->counters.p, C_COUNT * 4, cudaMemcpyDeviceToHost, ctx->L.stream));
        CU(ctx, cudaStreamSynchronize(ctx->L.stream));
        *is_changed = h[after_slot] != h[before_slot];
    }
    return RALA_B200_OK;
}

// graph.cpp:849-877 on demand: the final `internals` (alive at their own time, not kA/kB) and `overlaps`
// (both piles alive at the end, not kA/kB)
int materialize_final_lists(rala_b200_graph* g) {
    if (g->final_lists_ready || g->state < 3) return RALA_B200_OK;
    rala_b200_ctx* ctx = g->ctx;
    unsigned long long* status;
    uint32_t* ticket;
    List none{};
    g->scan_used = 0;
    CU(ctx, cudaMemsetAsync(g->scan_pool.p, 0, g->scan_pool_words * 8, ctx->L.stream));
    scan_state(g, g->cap, &status, &ticket);
    int inl_out = g->new_slot();
    CU(ctx, zero_counter(g, inl_out));
    launch_list_pass(ctx->L, 4 /*kFinalInt*/, g->inl[g->inl_cur].view, g->cnt() + g->slot_inl, g->cap, g->piles.as<uint2>(),
                     g->inl[g->inl_cur ^ 1].view, g->cnt() + inl_out, none, nullptr, nullptr, g->cap, g->dbuf.as<uint32_t>(),
                     g->n_piles, g->cnt() + g->final_time_base_slot, g->cnt(), status, ticket);
    g->inl_cur ^= 1;
    g->slot_inl = inl_out;
    scan_state(g, g->cap, &status, &ticket);
    int ovl_out = g->new_slot();
    CU(ctx, zero_counter(g, ovl_out));
    launch_list_pass(ctx->L, 3 /*kFinalOvl*/, g->ovl[g->ovl_cur].view, g->cnt() + g->slot_ovl, g->cap, g->piles.as<uint2>(),
                     g->ovl[g->ovl_cur ^ 1].view, g->cnt() + ovl_out, none, nullptr, nullptr, g->cap, nullptr, g->n_piles,
                     nullptr, g->cnt(), status, ticket);
    g->ovl_cur ^= 1;
    g->slot_ovl = ovl_out;
    CU(ctx, cudaGetLastError());
    g->final_lists_ready = true;
    return RALA_B200_OK;
}

// ---- graph.cpp:831-877 in two phases ---------------------------------------------------------------------
int phase_final_events(rala_b200_graph* g, const uint32_t* ovl_base /* device, nullable */,
                              const uint32_t* inl_base /* device */) {
    rala_b200_ctx* ctx = g->ctx;
    if (g->state != 2) return fail(ctx, RALA_B200_ERR_STATE, "finalize: classify first");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, begin_stage(g, ST_FINALIZE));
    CU(ctx, cudaMemcpyAsync(g->cnt() + C_EV_FIRST, g->cnt() + C_EV, 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    CU(ctx, cudaMemcpyAsync(g->cnt() + C_ROUNDS_FIRST, g->cnt() + C_ROUNDS, 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    CU(ctx, zero_counter(g, C_EV));
    CU(ctx, clear_victim_histogram(g));
    launch_classify_final(ctx->L, g->ovl[g->ovl_cur].view, g->cnt() + g->slot_ovl, ovl_base, g->inl[g->inl_cur].view,
                          g->cnt() + g->slot_inl, inl_base, g->cap, g->piles.as<uint2>(), g->events_view(), g->ev_cap,
                          resolve_bufs(g).vcursor, g->cnt());
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

static int phase_final_resolve(rala_b200_graph* g) {
    rala_b200_ctx* ctx = g->ctx;
    int rc = phase_resolve(g, false);
    if (rc) return rc;
    // The filtered `overlaps` / `internals` vectors of graph.cpp:867-877 are NOT materialised here: edge creation
    // (build) applies the same filter on the fly (both piles alive, type kAB/kBA), and nothing else on the path
    // reads them.  get_lists / counts materialise them on demand (materialize_final_lists).
    g->final_lists_ready = false;
    CU(ctx, end_stage(g, ST_FINALIZE));
    g->state = 3;
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_finalize(rala_b200_graph* g) {
    if (!g) return RALA_B200_ERR_ARG;
    g->final_time_base_slot = g->slot_ovl;   // internals are timed after every entry of `overlaps`
    int rc = phase_final_events(g, nullptr, g->cnt() + g->slot_ovl);
    if (!rc) rc = phase_final_resolve(g);
    return rc;
}

// graph.cpp:552-632
extern "C" int rala_b200_graph_build(rala_b200_graph* g) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state != 3) return fail(ctx, RALA_B200_ERR_STATE, "build: finalize first");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, begin_stage(g, ST_BUILD));
    CU(ctx, zero_counter(g, C_NODES, 4));   // C_NODES, C_ALIVE, C_DOVETAILS, C_EDGES
    CU(ctx, cudaMemsetAsync(g->cursor.p, 0, ((size_t) g->n_nodes_max + 8) * 4, ctx->L.stream));
    unsigned long long* status;
    uint32_t* ticket;
    scan_state(g, g->n_piles, &status, &ticket);
    launch_node_ids(ctx->L, g->piles.as<uint2>(), g->n_piles, g->seq_to_node.as<uint32_t>(), g->cnt(), status, ticket);
    scan_state(g, emit_scan_span(g->cap), &status, &ticket);
    launch_emit_edges(ctx->L, g->ovl[g->ovl_cur].view, g->cnt() + g->slot_ovl, g->cap, g->piles.as<uint2>(), g->graph_view(),
                      g->edge_cap, g->cnt(), status, ticket);
    // the edge list is final: rows into the caller's buffer on a forked stream, beside the CSR build and (inside a
    // whole run) the transitive pass; edges beyond the caller's capacity are not written (n_edges tells)
    if (g->out_edges && g->out_edges_cap && fork_side(ctx->L, 0)) {
        launch_pack_edges(ctx->L, ctx->L.side[0], g->graph_view(), g->edge_cap < g->out_edges_cap ? g->edge_cap : g->out_edges_cap,
                          g->cnt() + C_EDGES, g->out_edges);
        g->download_pending = true;
    }
    scan_state(g, (uint64_t) g->n_nodes_max + 1, &status, &ticket);
    launch_build_csr(ctx->L, g->graph_view(), g->n_nodes_max, g->edge_cap, g->cnt(), status, ticket, RB_OPT_RANK != 0);
    if (g->download_pending && !g->in_run) {
        join_side(ctx->L, 0);
        g->download_pending = false;
    }
    CU(ctx, cudaGetLastError());
    CU(ctx, end_stage(g, ST_BUILD));
    g->state = 4;
    return RALA_B200_OK;
}

static int run_transitive(rala_b200_graph* g) {
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, zero_counter(g, C_PAIRS, 2));   // C_PAIRS, C_HEAVY
    CU(ctx, zero_counter(g, C_HOP_LO, 2));
    CU(ctx, cudaMemsetAsync(g->work_counter.p, 0, 64, ctx->L.stream));
    CU(ctx, stage_event(g, g->ev_start[ST_K3_KERNELS]));
    launch_transitive(ctx->L, g->graph_view(), g->n_nodes_max, g->edge_cap, g->heavy_view(), g->work_counter.as<uint32_t>(),
                      g->cnt(), 0u, 0xFFFFFFFFu, g->world > 1 ? g->work_counter.as<uint32_t>() + 4 : nullptr);
    CU(ctx, end_stage(g, ST_K3_KERNELS));
    launch_finalize_marks(ctx->L, g->graph_view(), g->edge_cap, g->cnt(), g->out_marked, g->out_marked_cap);
    if (g->download_pending) {
        join_side(ctx->L, 0);
        g->download_pending = false;
    }
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

// graph.cpp:1281-1318
extern "C" int rala_b200_graph_transitive(rala_b200_graph* g) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 4) return fail(ctx, RALA_B200_ERR_STATE, "transitive: build first");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, begin_stage(g, ST_TRANSITIVE));
    int rc = run_transitive(g);
    if (rc) return rc;
    CU(ctx, end_stage(g, ST_TRANSITIVE));
    g->state = 5;
    return RALA_B200_OK;
}

static int run_eager(rala_b200_graph* g) {
    int rc = rala_b200_graph_classify(g);
    if (!rc) rc = rala_b200_graph_retrim(g);
    if (!rc) rc = rala_b200_graph_retrim_promote(g, nullptr);
    if (!rc) rc = rala_b200_graph_finalize(g);
    g->in_run = true;    // the transitive stage follows: it joins the forked edge download
    if (!rc) rc = rala_b200_graph_build(g);
    g->in_run = false;
    if (!rc) rc = rala_b200_graph_transitive(g);
    if (g->download_pending) {   // build forked, transitive failed before joining
        join_side(g->ctx->L, 0);
        g->download_pending = false;
    }
    return rc;
}

// Everything the sequence of stream operations of one run depends on.  The chain itself has no host
// synchronisation and every data-dependent size stays on the device, so for a fixed key the ~45 stream
// operations (27 kernels, memsets, small copies) are identical from run to run: the second run with a key is
// captured into a CUDA graph and later runs replay it with one launch.
struct RunKey {
    const void *rec, *recs, *piles, *piles_initial, *events, *segs, *dbuf, *edges, *col, *scan_pool, *tiles, *ovl0, *inl0, *hills;
    const void *out_edges, *out_marked;
    uint32_t n_rec, n_piles, n_hills, cap, ev_cap, edge_cap, heavy_cap, t0, out_edges_cap, out_marked_cap;
    int world, rank, coop_blocks;
    bool piles_fresh, skip_clean_retrim;
};

// host-side bookkeeping a run leaves behind (restored after a replay)
struct RunHostState {
    int ovl_cur, inl_cur, slot_ovl, slot_inl, next_slot, final_time_base_slot, state;
    bool final_lists_ready, piles_dirty, piles_fresh, promote_pending;
    uint32_t retrim_passes;
    size_t scan_used;
};

struct RunGraph {
    RunKey key{};
    bool seen = false;          // the key was run once (eagerly): capture on the next occurrence
    cudaGraphExec_t exec = nullptr;
    RunHostState after{};
    uint64_t launches = 0;      // kernels inside the graph
};

static RunKey run_key(const rala_b200_graph* g) {
    RunKey k;
    memset(&k, 0, sizeof(k));   // padding included: keys are compared with memcmp
    k.rec = g->rec.p; k.recs = g->recs.buf.p; k.piles = g->piles.p; k.piles_initial = g->piles_initial.p;
    k.events = g->events.p; k.segs = g->segs.p; k.dbuf = g->dbuf.p; k.edges = g->edges.p; k.col = g->col.p;
    k.scan_pool = g->scan_pool.p; k.tiles = g->tiles.p; k.ovl0 = g->ovl[0].buf.p; k.inl0 = g->inl[0].buf.p; k.hills = g->hills.p;
    k.n_rec = g->n_rec; k.n_piles = g->n_piles; k.n_hills = g->n_hills; k.cap = g->cap; k.ev_cap = g->ev_cap;
    k.edge_cap = g->edge_cap; k.heavy_cap = g->heavy_cap; k.t0 = g->t0;
    k.out_edges = g->out_edges; k.out_marked = g->out_marked; k.out_edges_cap = g->out_edges_cap; k.out_marked_cap = g->out_marked_cap;
    k.world = g->world; k.rank = g->rank; k.coop_blocks = g->ctx->coop_blocks;
    k.piles_fresh = g->piles_fresh; k.skip_clean_retrim = g->skip_clean_retrim;
    return k;
}

static RunHostState host_state(const rala_b200_graph* g) {
    return RunHostState{g->ovl_cur, g->inl_cur, g->slot_ovl, g->slot_inl, g->next_slot, g->final_time_base_slot, g->state,
                        g->final_lists_ready, g->piles_dirty, g->piles_fresh, g->promote_pending, g->retrim_passes, g->scan_used};
}

static void drop_run_graphs(rala_b200_graph* g) {
    if (!g->run_graphs) return;
    for (int i = 0; i < 2; ++i)
        if (g->run_graphs[i].exec) cudaGraphExecDestroy(g->run_graphs[i].exec);
    delete[] g->run_graphs;
    g->run_graphs = nullptr;
}

extern "C" int rala_b200_graph_use_cuda_graph(rala_b200_graph* g, int enabled) {
    if (!g) return RALA_B200_ERR_ARG;
    g->use_cuda_graph = enabled != 0;
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_run(rala_b200_graph* g) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (!g->use_cuda_graph) return run_eager(g);
    if (g->state < 1) return fail(ctx, RALA_B200_ERR_STATE, "run: set_overlaps and set_piles first");
    CU(ctx, cudaSetDevice(ctx->device));
    if (!g->run_graphs) g->run_graphs = new RunGraph[2];
    const RunKey key = run_key(g);
    RunGraph& R = g->run_graphs[key.piles_fresh ? 1 : 0];
    if (!R.seen || memcmp(&R.key, &key, sizeof(key)) != 0) {   // new shape: run it eagerly once (this also sizes every buffer)
        if (R.exec) cudaGraphExecDestroy(R.exec);
        R.exec = nullptr;
        R.key = key;
        R.seen = true;
        return run_eager(g);
    }
    if (!R.exec) {
        const uint64_t launches0 = ctx->L.count;
        cudaGraph_t graph = nullptr;
        CU(ctx, cudaStreamBeginCapture(ctx->L.stream, cudaStreamCaptureModeThreadLocal));
        g->capturing = true;
        const int rc = run_eager(g);
        g->capturing = false;
        const cudaError_t e = cudaStreamEndCapture(ctx->L.stream, &graph);
        if (rc || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            R.seen = false;
            ctx->L.count = launches0;
            if (rc) return rc;
            return fail(ctx, RALA_B200_ERR_CUDA, "run: stream capture failed: %s", cudaGetErrorString(e));
        }
        const cudaError_t ei = cudaGraphInstantiate(&R.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) {
            R.exec = nullptr;
            R.seen = false;
            return fail(ctx, RALA_B200_ERR_CUDA, "run: cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
        }
        R.launches = ctx->L.count - launches0;
        R.after = host_state(g);
        ctx->L.count = launches0;   // counted per replay below
    }
    CU(ctx, cudaGraphLaunch(R.exec, ctx->L.stream));
    ctx->L.count += R.launches;
    const RunHostState& a = R.after;
    g->ovl_cur = a.ovl_cur; g->inl_cur = a.inl_cur; g->slot_ovl = a.slot_ovl; g->slot_inl = a.slot_inl; g->next_slot = a.next_slot;
    g->final_time_base_slot = a.final_time_base_slot; g->state = a.state; g->final_lists_ready = a.final_lists_ready;
    g->piles_dirty = a.piles_dirty; g->piles_fresh = a.piles_fresh; g->promote_pending = a.promote_pending; g->retrim_passes = a.retrim_passes; g->scan_used = a.scan_used;
    for (int i = 0; i < RALA_B200_N_STAGES; ++i) g->ev_valid[i] = false;   // no stage timers inside a graph
    return RALA_B200_OK;
}

int read_counters(rala_b200_graph* g, uint32_t* h) {
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(h, g->counters.p, C_COUNT * 4, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    if (g->scan_pool_exhausted)
        return fail(ctx, RALA_B200_ERR_LIMIT, "the look-back status pool was exhausted inside a stage (%zu words): the results are void", g->scan_pool_words);
    if (h[C_OVERFLOW] || (g->state >= 2 && (h[g->slot_ovl] > g->cap || h[g->slot_inl] > g->cap)) || h[C_EV] > g->ev_cap || h[C_HILL] > g->cap || h[C_HEAVY] > g->heavy_cap)
        return fail(ctx, RALA_B200_ERR_LIMIT, "a device list overflowed its capacity (cap=%u events=%u hills=%u heavy=%u/%u)",
                    g->cap, h[C_EV], h[C_HILL], h[C_HEAVY], g->heavy_cap);
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_counts(rala_b200_graph* g, rala_b200_counts_t* out) {
    if (!g || !out) return RALA_B200_ERR_ARG;
    uint32_t h[C_COUNT];
    int rc = materialize_final_lists(g);
    if (!rc) rc = read_counters(g, h);
    if (rc) return rc;
    memset(out, 0, sizeof(*out));
    out->n_records = g->n_rec;
    out->n_overlaps = g->state >= 2 ? h[g->slot_ovl] : 0;
    out->n_internals = g->state >= 2 ? h[g->slot_inl] : 0;
    out->n_candidates = g->state >= 3 ? h[C_EV_FIRST] : h[C_EV];
    out->n_rounds = g->state >= 3 ? h[C_ROUNDS_FIRST] : h[C_ROUNDS];
    out->n_final_candidates = g->state >= 3 ? h[C_EV] : 0;
    out->n_final_rounds = g->state >= 3 ? h[C_ROUNDS] : 0;
    out->n_piles = g->n_piles;
    out->n_alive_piles = h[C_ALIVE];
    out->n_nodes = h[C_NODES];
    out->n_edges = h[C_EDGES];
    out->n_two_hop = (uint64_t) h[C_HOP_LO] | ((uint64_t) h[C_HOP_HI] << 32);
    out->n_transitive_pairs = h[C_PAIRS];
    out->n_heavy_items = h[C_HEAVY];
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_get_hill_coverage(rala_b200_graph* g, uint32_t* cov_out) {
    if (!g || (g->n_hills && !cov_out)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (g->n_hills)
        CU(ctx, cudaMemcpyAsync(cov_out, g->hills.as<uint32_t>() + 3 * (size_t) g->n_hills, (size_t) g->n_hills * 4,
                                cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_get_piles(rala_b200_graph* g, rala_pile_t* piles_out) {
    if (!g || !piles_out) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    launch_unpack_piles(ctx->L, g->piles.as<uint2>(), g->piles_raw.as<uint2>(), g->n_piles);
    CU(ctx, cudaMemcpyAsync(piles_out, g->piles_raw.p, (size_t) g->n_piles * 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_get_connections(rala_b200_graph* g, uint32_t* ab_out) {
    if (!g || !ab_out) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 2) return fail(ctx, RALA_B200_ERR_STATE, "get_connections: classify first");
    uint32_t h[C_COUNT];
    int rc = read_counters(g, h);
    if (rc) return rc;
    uint32_t n = h[g->slot_ovl];
    if (!n) return RALA_B200_OK;
    DevBuf tmp;
    CU(ctx, tmp.reserve((size_t) n * 8));
    launch_list_connections(ctx->L, g->ovl[g->ovl_cur].view, g->cnt() + g->slot_ovl, g->cap, tmp.as<uint32_t>());
    CU(ctx, cudaMemcpyAsync(ab_out, tmp.p, (size_t) n * 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    tmp.release();
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_get_lists(rala_b200_graph* g, rala_ovl_t* overlaps_out, rala_ovl_t* internals_out) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 2) return fail(ctx, RALA_B200_ERR_STATE, "get_lists: classify first");
    uint32_t h[C_COUNT];
    int rc = materialize_final_lists(g);
    if (!rc) rc = read_counters(g, h);
    if (rc) return rc;
    DevBuf tmp;
    for (int which = 0; which < 2; ++which) {
        rala_ovl_t* out = which ? internals_out : overlaps_out;
        uint32_t n = which ? h[g->slot_inl] : h[g->slot_ovl];
        if (!out || !n) continue;
        CU(ctx, tmp.reserve((size_t) n * 28));
        launch_list_to_aos(ctx->L, which ? g->inl[g->inl_cur].view : g->ovl[g->ovl_cur].view,
                           g->cnt() + (which ? g->slot_inl : g->slot_ovl), g->cap, tmp.as<uint32_t>());
        CU(ctx, cudaMemcpyAsync(out, tmp.p, (size_t) n * 28, cudaMemcpyDeviceToHost, ctx->L.stream));
        CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    }
    tmp.release();
    return RALA_B200_OK;
}

// graph.cpp:523 / 882-1054: the host filtered `overlaps` (Graph::preprocess(overlaps, sensitive_path) only ever drops
// entries); the device list is replaced by what is left, in the same order.
extern "C" int rala_b200_graph_set_kept_overlaps(rala_b200_graph* g, const rala_ovl_t* kept, uint64_t n) {
    if (!g || (n && !kept)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state != 3) return fail(ctx, RALA_B200_ERR_STATE, "set_kept_overlaps: finalize first");
    if (n > g->cap) return fail(ctx, RALA_B200_ERR_LIMIT, "set_kept_overlaps: %llu entries exceed the list capacity %u", (unsigned long long) n, g->cap);
    CU(ctx, cudaSetDevice(ctx->device));
    int rc = materialize_final_lists(g);   // `internals` keeps its final form
    if (rc) return rc;
    DevBuf tmp;
    CU(ctx, tmp.reserve((size_t) n * 28 + 16));
    if (n) CU(ctx, cudaMemcpyAsync(tmp.p, kept, (size_t) n * 28, cudaMemcpyHostToDevice, ctx->L.stream));
    launch_aos_to_list(ctx->L, tmp.as<uint32_t>(), (uint32_t) n, g->ovl[g->ovl_cur].view);
    const uint32_t n32 = (uint32_t) n;
    CU(ctx, cudaMemcpyAsync(g->cnt() + g->slot_ovl, &n32, 4, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    tmp.release();
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_get_seq_to_node(rala_b200_graph* g, uint32_t* out) {
    if (!g || !out) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 4) return fail(ctx, RALA_B200_ERR_STATE, "get_seq_to_node: build first");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(out, g->seq_to_node.p, (size_t) g->n_piles * 4, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_get_edges(rala_b200_graph* g, rala_edge_t* out) {
    if (!g || !out) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 4) return fail(ctx, RALA_B200_ERR_STATE, "get_edges: build first");
    uint32_t h[C_COUNT];
    int rc = read_counters(g, h);
    if (rc) return rc;
    uint32_t n = h[C_EDGES];
    if (!n) return RALA_B200_OK;
    // columns -> rows on the device, then one contiguous copy
    CU(ctx, g->edges_aos.reserve((size_t) g->edge_cap * 12));
    launch_pack_edges(ctx->L, ctx->L.stream, g->graph_view(), g->edge_cap, g->cnt() + C_EDGES, g->edges_aos.as<uint32_t>());
    CU(ctx, cudaMemcpyAsync(out, g->edges_aos.p, (size_t) n * 12, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_get_marked(rala_b200_graph* g, uint8_t* out) {
    if (!g || !out) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 5) return fail(ctx, RALA_B200_ERR_STATE, "get_marked: transitive first");
    uint32_t h[C_COUNT];
    int rc = read_counters(g, h);
    if (rc) return rc;
    if (h[C_EDGES]) CU(ctx, cudaMemcpyAsync(out, g->marked.p, h[C_EDGES], cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    return RALA_B200_OK;
}

// suffix_edges_ / prefix_edges_ of every node as the reference keeps them (ascending edge id), optionally after the removal of
// the marked edges: graph.cpp:603-606 / 622-625 and 2118-2151
extern "C" int rala_b200_graph_get_adjacency(rala_b200_graph* g, int which, int skip_marked, uint32_t* off_out, uint32_t* ids_out,
                                             uint64_t* n_ids_out) {
    if (!g || !off_out || !ids_out || !n_ids_out || which < 0 || which > 1) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 4) return fail(ctx, RALA_B200_ERR_STATE, "get_adjacency: build first");
    if (skip_marked && g->state < 5) return fail(ctx, RALA_B200_ERR_STATE, "get_adjacency: transitive first (or pass skip_marked = 0)");
    uint32_t h[C_COUNT];
    int rc = read_counters(g, h);
    if (rc) return rc;
    const uint32_t n_nodes = h[C_NODES], n_edges = h[C_EDGES];
    *n_ids_out = 0;
    CU(ctx, cudaSetDevice(ctx->device));
    DevBuf deg, rp, tmp, sorted;
    CU(ctx, deg.reserve(((size_t) n_nodes + 8) * 4));
    CU(ctx, rp.reserve(((size_t) n_nodes + 8) * 4));
    CU(ctx, tmp.reserve(((size_t) n_edges + 8) * 4));
    CU(ctx, sorted.reserve(((size_t) n_edges + 8) * 4));
    GraphArrays ga = g->graph_view();
    g->scan_used = 0;
    CU(ctx, cudaMemsetAsync(g->scan_pool.p, 0, g->scan_pool_words * 8, ctx->L.stream));
    unsigned long long* status;
    uint32_t* ticket;
    scan_state(g, (uint64_t) n_nodes + 1, &status, &ticket);
    launch_adjacency_view(ctx->L, which ? ga.dst : ga.src, skip_marked ? ga.marked : nullptr, g->cnt() + C_EDGES, g->edge_cap, n_nodes,
                          deg.as<uint32_t>(), rp.as<uint32_t>(), tmp.as<uint32_t>(), sorted.as<uint32_t>(), status, ticket);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(off_out, rp.p, ((size_t) n_nodes + 1) * 4, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    const uint32_t total = off_out[n_nodes];
    if (total) CU(ctx, cudaMemcpyAsync(ids_out, sorted.p, (size_t) total * 4, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    *n_ids_out = total;
    deg.release(); rp.release(); tmp.release(); sorted.release();
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_stage_ms(rala_b200_graph* g, float* ms_out) {
    if (!g || !ms_out) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    for (int i = 0; i < RALA_B200_N_STAGES; ++i) {
        ms_out[i] = 0.f;
        if (g->ev_valid[i]) CU(ctx, cudaEventElapsedTime(&ms_out[i], g->ev_start[i], g->ev_stop[i]));
    }
    return RALA_B200_OK;
}

// ---------------------------------------------------------------------------------------------
// stateless stages
// ---------------------------------------------------------------------------------------------
extern "C" int rala_b200_trim_classify(rala_b200_ctx* ctx, rala_ovl_t* ovl, uint64_t n, const rala_pile_t* piles,
                                       uint32_t n_piles, uint8_t* type_out) {
    if (!ctx || (n && (!ovl || !type_out)) || (n_piles && !piles)) return RALA_B200_ERR_ARG;
    if (n >= (1ull << 31)) return fail(ctx, RALA_B200_ERR_LIMIT, "too many records");
    if (n == 0) return RALA_B200_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    DevBuf rec, praw, ppacked, types;
    CU(ctx, rec.reserve((size_t) n * 28));
    CU(ctx, praw.reserve((size_t) n_piles * 8 + 16));
    CU(ctx, ppacked.reserve((size_t) n_piles * 8 + 16));
    CU(ctx, types.reserve(n));
    CU(ctx, cudaMemcpyAsync(rec.p, ovl, (size_t) n * 28, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx, cudaMemcpyAsync(praw.p, piles, (size_t) n_piles * 8, cudaMemcpyHostToDevice, ctx->L.stream));
    launch_pack_piles(ctx->L, praw.as<uint2>(), nullptr, ppacked.as<uint2>(), n_piles);
    launch_trim_classify_aos(ctx->L, rec.as<uint32_t>(), (uint32_t) n, ppacked.as<uint2>(), n_piles, types.as<uint8_t>());
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(ovl, rec.p, (size_t) n * 28, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaMemcpyAsync(type_out, types.p, n, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    rec.release(); praw.release(); ppacked.release(); types.release();
    return RALA_B200_OK;
}

extern "C" int rala_b200_transitive_reduce(rala_b200_ctx* ctx, uint32_t n_nodes, uint64_t n_edges, const rala_edge_t* edges,
                                           uint8_t* marked_out, uint64_t* n_pairs) {
    if (!ctx || (n_edges && (!edges || !marked_out))) return RALA_B200_ERR_ARG;
    if (n_edges >= (1ull << 31) || (n_edges & 1)) return fail(ctx, RALA_B200_ERR_ARG, "edge count must be even and < 2^31 (pair(e) = e ^ 1)");
    if (n_pairs) *n_pairs = 0;
    if (n_edges == 0) return RALA_B200_OK;
    for (uint64_t i = 0; i < n_edges; ++i) {
        if (edges[i].src >= n_nodes || edges[i].dst >= n_nodes) return fail(ctx, RALA_B200_ERR_ARG, "edge %llu references a node >= n_nodes", (unsigned long long) i);
    }
    CU(ctx, cudaSetDevice(ctx->device));
    // a throw-away session provides the buffers
    rala_b200_graph* g = nullptr;
    int rc = rala_b200_graph_create(ctx, &g);
    if (rc) return rc;
    g->n_piles = (n_nodes + 1) / 2;
    g->n_nodes_max = n_nodes;
    g->edge_cap = (uint32_t) n_edges;
    g->heavy_cap = g->edge_cap / 16 + 4096;
    DevBuf rows;
    cudaError_t e = rows.reserve((size_t) n_edges * 12);
    if (e == cudaSuccess) e = g->edges.reserve(align_up((size_t) g->edge_cap * 4, 256) * 3);
    if (e == cudaSuccess) e = g->col.reserve((size_t) g->edge_cap * 8);
    if (e == cudaSuccess) e = g->col_eid.reserve((size_t) g->edge_cap * 4);
    if (e == cudaSuccess) e = g->T.reserve(align_up(g->edge_cap, 256));
    if (e == cudaSuccess) e = g->marked.reserve(align_up(g->edge_cap, 256));
    if (e == cudaSuccess) e = g->heavy.reserve(align_up((size_t) g->heavy_cap * 4, 256) * 3);
    if (e == cudaSuccess) e = g->row_ptr.reserve(((size_t) n_nodes + 8) * 4);
    if (e == cudaSuccess) e = g->cursor.reserve(((size_t) n_nodes + 8) * 4);
    g->scan_pool_words = 4 * (tiles_of((uint64_t) n_nodes + 1) + 8);
    if (e == cudaSuccess) e = g->scan_pool.reserve(g->scan_pool_words * 8);
    if (e != cudaSuccess) {
        rows.release();
        rala_b200_graph_destroy(g);
        return fail(ctx, RALA_B200_ERR_CUDA, "transitive_reduce: %s", cudaGetErrorString(e));
    }
    auto cleanup = [&](int code) {
        rows.release();
        rala_b200_graph_destroy(g);
        return code;
    };
    GraphArrays ga = g->graph_view();
    cudaStream_t s = ctx->L.stream;
#define CUX(call)                                                                                      \
    do {                                                                                               \
        cudaError_t err__ = (call);                                                                    \
        if (err__ != cudaSuccess) return cleanup(fail(ctx, RALA_B200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(err__))); \
    } while (0)
    CUX(cudaMemcpyAsync(rows.p, edges, (size_t) n_edges * 12, cudaMemcpyHostToDevice, s));
    CUX(cudaMemcpy2DAsync(ga.src, 4, rows.as<char>(), 12, 4, n_edges, cudaMemcpyDeviceToDevice, s));
    CUX(cudaMemcpy2DAsync(ga.dst, 4, rows.as<char>() + 4, 12, 4, n_edges, cudaMemcpyDeviceToDevice, s));
    CUX(cudaMemcpy2DAsync(ga.len, 4, rows.as<char>() + 8, 12, 4, n_edges, cudaMemcpyDeviceToDevice, s));
    uint32_t hc[C_COUNT];
    memset(hc, 0, sizeof(hc));
    hc[C_NODES] = n_nodes;
    hc[C_EDGES] = (uint32_t) n_edges;
    CUX(cudaMemcpyAsync(g->counters.p, hc, sizeof(hc), cudaMemcpyHostToDevice, s));
    CUX(cudaMemsetAsync(g->cursor.p, 0, ((size_t) n_nodes + 8) * 4, s));
    CUX(cudaMemsetAsync(g->scan_pool.p, 0, g->scan_pool_words * 8, s));
    launch_degree_hist(ctx->L, ga.src, g->cnt() + C_EDGES, g->edge_cap, ga.cursor);
    unsigned long long* status;
    uint32_t* ticket;
    scan_state(g, (uint64_t) n_nodes + 1, &status, &ticket);
    launch_build_csr(ctx->L, ga, n_nodes, g->edge_cap, g->cnt(), status, ticket);
    rc = run_transitive(g);
    if (rc) return cleanup(rc);
    uint32_t h[C_COUNT];
    CUX(cudaMemcpyAsync(marked_out, g->marked.p, n_edges, cudaMemcpyDeviceToHost, s));
    CUX(cudaMemcpyAsync(h, g->counters.p, sizeof(h), cudaMemcpyDeviceToHost, s));
    CUX(cudaStreamSynchronize(s));
#undef CUX
    if (h[C_OVERFLOW] || h[C_HEAVY] > g->heavy_cap) return cleanup(fail(ctx, RALA_B200_ERR_LIMIT, "heavy work list overflow"));
    if (n_pairs) *n_pairs = h[C_PAIRS];
    return cleanup(RALA_B200_OK);
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU phases (one process per GPU).  The records are sharded by contiguous file range, the pile
// table is replicated; the collectives between the phases belong to the caller (NCCL through
// torch.distributed in rala_b200/multi.py), which passes DEVICE pointers of its exchange buffers:
//   phase_events -> [all-gather events] -> import_events -> phase_resolve(1) -> phase_survivors
//   -> [all-gather list counts] -> phase_final_events -> [all-gather events] -> import_events
//   -> phase_final_resolve -> phase_emit_edges -> [all-gather edges] -> import_edges -> phase_csr
//   -> phase_transitive(rank, world) -> [all-reduce(max) marks] -> phase_marks
// Exchange blocks are three columns of `stride` words: events v | c | t, edges src | dst | len.
// ---------------------------------------------------------------------------------------------
extern "C" int rala_b200_graph_set_shard(rala_b200_graph* g, uint32_t t0, int rank, int world) {
    if (!g || world < 1 || rank < 0 || rank >= world) return RALA_B200_ERR_ARG;
    g->t0 = t0;
    g->rank = rank;
    g->world = world;
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_phase_events(rala_b200_graph* g) { return g ? phase_events(g) : RALA_B200_ERR_ARG; }

static int read_counter(rala_b200_graph* g, int slot, uint32_t* out) {
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(out, g->cnt() + slot, 4, cudaMemcpyDeviceToHost, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_events_count(rala_b200_graph* g, uint32_t* n) {
    if (!g || !n) return RALA_B200_ERR_ARG;
    int rc = read_counter(g, C_EV, n);
    if (!rc && *n > g->ev_cap) return fail(g->ctx, RALA_B200_ERR_LIMIT, "event list overflow (%u > %u)", *n, g->ev_cap);
    return rc;
}

extern "C" int rala_b200_graph_export_events(rala_b200_graph* g, uint32_t* d_cols, uint32_t stride, uint32_t n) {
    if (!g || (n && !d_cols) || n > stride) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    Events ev = g->events_view();
    if (n) {
        CU(ctx, cudaMemcpyAsync(d_cols, ev.v, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
        CU(ctx, cudaMemcpyAsync(d_cols + stride, ev.c, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
        CU(ctx, cudaMemcpyAsync(d_cols + 2 * (size_t) stride, ev.t, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    }
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_import_events(rala_b200_graph* g, const uint32_t* d_cols, uint32_t stride, uint32_t n,
                                             uint32_t offset, uint32_t total) {
    if (!g || (n && !d_cols) || (uint64_t) offset + n > total) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (offset == 0) {
        int rc = reserve_events(g, total);   // the local events were exported before the first import
        if (rc) return rc;
    }
    Events ev = g->events_view();
    if (n) {
        CU(ctx, cudaMemcpyAsync(ev.v + offset, d_cols, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
        CU(ctx, cudaMemcpyAsync(ev.c + offset, d_cols + stride, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
        CU(ctx, cudaMemcpyAsync(ev.t + offset, d_cols + 2 * (size_t) stride, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    }
    if ((uint64_t) offset + n == total) {
        CU(ctx, cudaMemcpyAsync(g->cnt() + C_EV, &total, 4, cudaMemcpyHostToDevice, ctx->L.stream));
        CU(ctx, cudaStreamSynchronize(ctx->L.stream));   // `total` lives on this stack frame
        CU(ctx, clear_victim_histogram(g));
        launch_events_hist(ctx->L, g->events_view(), g->cnt() + C_EV, g->ev_cap, resolve_bufs(g).vcursor);
        CU(ctx, cudaGetLastError());
    }
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_phase_resolve(rala_b200_graph* g, int first_pass) {
    if (!g) return RALA_B200_ERR_ARG;
    return first_pass ? phase_resolve(g, true) : phase_final_resolve(g);
}

extern "C" int rala_b200_graph_phase_survivors(rala_b200_graph* g) { return g ? phase_survivors(g) : RALA_B200_ERR_ARG; }

extern "C" int rala_b200_graph_list_counts(rala_b200_graph* g, uint32_t* n_ovl, uint32_t* n_int) {
    if (!g || !n_ovl || !n_int) return RALA_B200_ERR_ARG;
    if (g->state < 2) return fail(g->ctx, RALA_B200_ERR_STATE, "list_counts: classify first");
    uint32_t h[C_COUNT];
    int rc = read_counters(g, h);
    if (rc) return rc;
    *n_ovl = h[g->slot_ovl];
    *n_int = h[g->slot_inl];
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_phase_final_events(rala_b200_graph* g, uint32_t ovl_base, uint32_t int_base) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    uint32_t bases[2] = {ovl_base, int_base};
    CU(ctx, cudaMemcpyAsync(g->cnt() + C_TBASE_OVL, bases, 8, cudaMemcpyHostToDevice, ctx->L.stream));
    CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    g->final_time_base_slot = C_TBASE_INL;
    return phase_final_events(g, g->cnt() + C_TBASE_OVL, g->cnt() + C_TBASE_INL);
}

// node ids (replicated) + the edges of the LOCAL overlaps, in local list order, with global node ids
extern "C" int rala_b200_graph_phase_emit_edges(rala_b200_graph* g, uint32_t* n_local_edges) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state != 3) return fail(ctx, RALA_B200_ERR_STATE, "emit_edges: finalize first");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, begin_stage(g, ST_BUILD));
    CU(ctx, zero_counter(g, C_NODES, 4));
    CU(ctx, cudaMemsetAsync(g->cursor.p, 0, ((size_t) g->n_nodes_max + 8) * 4, ctx->L.stream));
    unsigned long long* status;
    uint32_t* ticket;
    scan_state(g, g->n_piles, &status, &ticket);
    launch_node_ids(ctx->L, g->piles.as<uint2>(), g->n_piles, g->seq_to_node.as<uint32_t>(), g->cnt(), status, ticket);
    scan_state(g, emit_scan_span(g->cap), &status, &ticket);
    launch_emit_edges(ctx->L, g->ovl[g->ovl_cur].view, g->cnt() + g->slot_ovl, g->cap, g->piles.as<uint2>(), g->graph_view(),
                      g->edge_cap, g->cnt(), status, ticket);
    CU(ctx, cudaGetLastError());
    return n_local_edges ? read_counter(g, C_EDGES, n_local_edges) : RALA_B200_OK;   // NULL: no host synchronisation
}

extern "C" int rala_b200_graph_export_edges(rala_b200_graph* g, uint32_t* d_cols, uint32_t stride, uint32_t n) {
    if (!g || (n && !d_cols) || n > stride) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    GraphArrays ga = g->graph_view();
    if (n) {
        CU(ctx, cudaMemcpyAsync(d_cols, ga.src, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
        CU(ctx, cudaMemcpyAsync(d_cols + stride, ga.dst, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
        CU(ctx, cudaMemcpyAsync(d_cols + 2 * (size_t) stride, ga.len, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    }
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_import_edges(rala_b200_graph* g, const uint32_t* d_cols, uint32_t stride, uint32_t n,
                                            uint32_t offset, uint32_t total) {
    if (!g || (n && !d_cols) || (uint64_t) offset + n > total) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (offset == 0) {
        int rc = reserve_edges(g, total);   // the local edges were exported before the first import
        if (rc) return rc;
    }
    GraphArrays ga = g->graph_view();
    if (n) {
        CU(ctx, cudaMemcpyAsync(ga.src + offset, d_cols, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
        CU(ctx, cudaMemcpyAsync(ga.dst + offset, d_cols + stride, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
        CU(ctx, cudaMemcpyAsync(ga.len + offset, d_cols + 2 * (size_t) stride, (size_t) n * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    }
    if ((uint64_t) offset + n == total) {
        CU(ctx, cudaMemcpyAsync(g->cnt() + C_EDGES, &total, 4, cudaMemcpyHostToDevice, ctx->L.stream));
        CU(ctx, cudaStreamSynchronize(ctx->L.stream));
    }
    return RALA_B200_OK;
}

// ---- capacity-bounded exchange: counts travel inside the blocks, nothing is read back by the host ------------------------
extern "C" int rala_b200_graph_export_padded(rala_b200_graph* g, int kind, uint32_t* d_block, uint32_t cap) {
    if (!g || !d_block || cap == 0 || (kind != 0 && kind != 1)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (kind == 0) {
        Events ev = g->events_view();
        launch_export_padded(ctx->L, ev.v, ev.c, ev.t, g->cnt() + C_EV, g->ev_cap, cap, d_block);
    } else {
        GraphArrays ga = g->graph_view();
        launch_export_padded(ctx->L, ga.src, ga.dst, ga.len, g->cnt() + C_EDGES, g->edge_cap, cap, d_block);
    }
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_import_gathered(rala_b200_graph* g, int kind, const uint32_t* d_gathered, uint32_t cap, int world) {
    if (!g || !d_gathered || cap == 0 || world < 1 || (kind != 0 && kind != 1)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    const uint64_t total_cap = (uint64_t) cap * world;
    if (total_cap >= (1ull << 31)) return fail(ctx, RALA_B200_ERR_LIMIT, "exchange capacity too large");
    if (kind == 0) {
        int rc = reserve_events(g, (uint32_t) total_cap);   // the local events were exported before
        if (rc) return rc;
        Events ev = g->events_view();
        launch_import_gathered(ctx->L, d_gathered, cap, (uint32_t) world, ev.v, ev.c, ev.t, g->ev_cap, g->cnt() + C_EV, g->cnt() + C_OVERFLOW);
        CU(ctx, clear_victim_histogram(g));
        launch_events_hist(ctx->L, g->events_view(), g->cnt() + C_EV, g->ev_cap, resolve_bufs(g).vcursor);
    } else {
        int rc = reserve_edges(g, (uint32_t) total_cap);     // the local edges were exported before
        if (rc) return rc;
        GraphArrays ga = g->graph_view();
        launch_import_gathered(ctx->L, d_gathered, cap, (uint32_t) world, ga.src, ga.dst, ga.len, g->edge_cap, g->cnt() + C_EDGES, g->cnt() + C_OVERFLOW);
    }
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

extern "C" uint64_t rala_b200_exchange_block_words(int kind, uint32_t cap) {
    (void) kind;
    return 3ull * cap + 4ull;
}

extern "C" int rala_b200_graph_export_list_counts(rala_b200_graph* g, uint32_t* d_pair) {
    if (!g || !d_pair) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 2) return fail(ctx, RALA_B200_ERR_STATE, "export_list_counts: classify first");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(d_pair, g->cnt() + g->slot_ovl, 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    CU(ctx, cudaMemcpyAsync(d_pair + 1, g->cnt() + g->slot_inl, 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_phase_final_events_gathered(rala_b200_graph* g, const uint32_t* d_counts, int world) {
    if (!g || !d_counts || world < 1) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    launch_time_bases(ctx->L, d_counts, (uint32_t) g->rank, (uint32_t) world, g->cnt() + C_TBASE_OVL);
    g->final_time_base_slot = C_TBASE_INL;
    return phase_final_events(g, g->cnt() + C_TBASE_OVL, g->cnt() + C_TBASE_INL);
}

// CSR over ALL edges (replicated on every rank)
extern "C" int rala_b200_graph_phase_csr(rala_b200_graph* g) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemsetAsync(g->cursor.p, 0, ((size_t) g->n_nodes_max + 8) * 4, ctx->L.stream));
    GraphArrays ga = g->graph_view();
    launch_degree_hist(ctx->L, ga.src, g->cnt() + C_EDGES, g->edge_cap, ga.cursor);
    unsigned long long* status;
    uint32_t* ticket;
    scan_state(g, (uint64_t) g->n_nodes_max + 1, &status, &ticket);
    launch_build_csr(ctx->L, ga, g->n_nodes_max, g->edge_cap, g->cnt(), status, ticket);
    CU(ctx, cudaGetLastError());
    CU(ctx, end_stage(g, ST_BUILD));
    g->state = 4;
    return RALA_B200_OK;
}

// T(e) for the candidate edges whose source node this rank owns
extern "C" int rala_b200_graph_phase_transitive(rala_b200_graph* g) {
    if (!g) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    if (g->state < 4) return fail(ctx, RALA_B200_ERR_STATE, "transitive: build first");
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, begin_stage(g, ST_TRANSITIVE));
    CU(ctx, zero_counter(g, C_PAIRS, 2));
    CU(ctx, zero_counter(g, C_HOP_LO, 2));
    CU(ctx, cudaMemsetAsync(g->work_counter.p, 0, 64, ctx->L.stream));
    launch_node_range(ctx->L, g->graph_view(), g->cnt(), (uint32_t) g->rank, (uint32_t) g->world, g->work_counter.as<uint32_t>() + 4);
    CU(ctx, stage_event(g, g->ev_start[ST_K3_KERNELS]));
    launch_transitive(ctx->L, g->graph_view(), g->n_nodes_max, g->edge_cap, g->heavy_view(), g->work_counter.as<uint32_t>(),
                      g->cnt(), 0u, 0xFFFFFFFFu, g->work_counter.as<uint32_t>() + 4);
    CU(ctx, end_stage(g, ST_K3_KERNELS));
    CU(ctx, cudaGetLastError());
    return RALA_B200_OK;
}

extern "C" int rala_b200_graph_export_marks(rala_b200_graph* g, uint8_t* d_T, uint32_t n) {
    if (!g || (n && !d_T)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (n) CU(ctx, cudaMemcpyAsync(d_T, g->T.p, n, cudaMemcpyDeviceToDevice, ctx->L.stream));
    return RALA_B200_OK;
}

// merged T in, marked(e) = T(e) | T(e^1) out
extern "C" int rala_b200_graph_phase_marks(rala_b200_graph* g, const uint8_t* d_T, uint32_t n) {
    if (!g || (n && !d_T)) return RALA_B200_ERR_ARG;
    rala_b200_ctx* ctx = g->ctx;
    CU(ctx, cudaSetDevice(ctx->device));
    if (n) CU(ctx, cudaMemcpyAsync(g->T.p, d_T, n, cudaMemcpyDeviceToDevice, ctx->L.stream));
    launch_finalize_marks(ctx->L, g->graph_view(), g->edge_cap, g->cnt(), nullptr, 0u);
    CU(ctx, cudaGetLastError());
    CU(ctx, end_stage(g, ST_TRANSITIVE));
    g->state = 5;
    return RALA_B200_OK;
}
