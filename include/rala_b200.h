/* rala_b200.h — C ABI of the B200-native assembly-graph hot path (librala_b200.so).
 *
 * The reference (rvaser/rala) has no plugin / FFI interface: `rala` is one executable and the state
 * of rala::Graph is private.  The boundary is therefore cut at the two member functions its CLI
 * already calls:
 *     Graph::construct               /root/reference/src/graph.cpp:427-640   (main.cpp:74)
 *     Graph::remove_transitive_edges /root/reference/src/graph.cpp:1281-1335 (via simplify, graph.cpp:646)
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows the
 * reference-side patch (plain C++ calls; no binding layer is needed since the reference is C++).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no CUDA, torch or C++ types in any signature.
 *   - every function returns 0 on success or a negative rala_b200_status; the message is available
 *     from rala_b200_last_error().  The host shim keeps the reference's own convention on top of
 *     that: fprintf(stderr, "[rala::Graph::construct] error: %s!\n", ...); exit(1);  (cf. graph.cpp:418-421)
 *   - called from one host thread at a time per context (the reference drives this path from its
 *     main thread only, SURVEY.md finding 1).
 *   - there is NO CPU fallback: without a CUDA device rala_b200_create() fails.
 *
 * Limits: ids, coordinates and counts are 32-bit; read lengths < 2^30;
 * overlap records per context < 2^31.
 */
#ifndef RALA_B200_H
#define RALA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RALA_B200_ABI_VERSION 2

/* One PAF/MHAP overlap after name->id translation; replaces the numeric members of rala::Overlap
 * (overlap.hpp:104-116).  Records are passed in FILE ORDER: the position of a record is its
 * processing time in the reference's order-dependent containment removal (graph.cpp:448-480). */
typedef struct {
    uint32_t a_id, b_id;       /* sequence ids (Overlap::a_id_/b_id_ after transmute, overlap.cpp:36-82) */
    uint32_t a_begin, a_end;   /* on a's forward strand */
    uint32_t b_begin, b_end;   /* on b's forward strand */
    uint32_t flags;            /* bit0: orientation_ (overlap.cpp:18, 29); bit1: record is invalid —
                                  is_valid_overlap_ false (graph.cpp:450) or a name that is not in name_to_id_ */
} rala_ovl_t;
#define RALA_OVL_RC       1u
#define RALA_OVL_INVALID  2u

/* Valid region [begin, end) of one read's pile (Pile::begin()/end(), pile.hpp:33-42).
 * end == 0 means the pile is dead (piles_[i] == nullptr in the reference). */
typedef struct { uint32_t begin, end; } rala_pile_t;

/* per-pile flag byte */
#define RALA_PILE_HAS_HILL    1u   /* Pile::has_chimeric_hill()   (pile.hpp:107) */
#define RALA_PILE_HAS_REGION  2u   /* Pile::has_chimeric_region() (pile.hpp:121) */

/* One chimeric hill (Pile::chimeric_hills_[k]); rows grouped by ascending pile id, in the pile's own order. */
typedef struct { uint32_t pile, begin, end; } rala_hill_t;

/* One directed edge; row index = edge id, pair(e) = e ^ 1, pair(node) = node ^ 1 (graph.cpp:553-632). */
typedef struct { uint32_t src, dst, len; } rala_edge_t;

/* Overlap::type() values (overlap.hpp:27-33) + 255 for "rejected by trim / dead pile / invalid". */
enum { RALA_KX = 0, RALA_KA = 1, RALA_KB = 2, RALA_KAB = 3, RALA_KBA = 4, RALA_REJECTED = 255 };

typedef enum {
    RALA_B200_OK = 0,
    RALA_B200_ERR_CUDA = -1,        /* a CUDA call failed (message has the CUDA error string) */
    RALA_B200_ERR_ARG = -2,         /* bad argument */
    RALA_B200_ERR_STATE = -3,       /* stage called out of order */
    RALA_B200_ERR_NO_DEVICE = -4,   /* no sm_100 device: there is no CPU fallback */
    RALA_B200_ERR_LIMIT = -5        /* a documented limit was exceeded */
} rala_b200_status;

typedef struct rala_b200_ctx rala_b200_ctx;      /* one CUDA device + stream + scratch arena */
typedef struct rala_b200_graph rala_b200_graph;  /* device-resident state of one construct + reduce */

int rala_b200_abi_version(void);
int rala_b200_create(rala_b200_ctx** out, int device);
void rala_b200_destroy(rala_b200_ctx* ctx);
const char* rala_b200_last_error(const rala_b200_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t rala_b200_launch_count(const rala_b200_ctx* ctx);
/* CUDA events on the context's stream: record(0) ... record(1), then elapsed_ms waits for event 1. */
int rala_b200_event_record(rala_b200_ctx* ctx, int which);
int rala_b200_event_elapsed_ms(rala_b200_ctx* ctx, float* ms);
int rala_b200_synchronize(rala_b200_ctx* ctx);

/* ------------------------------------------------------------------------------------------------
 * Stateless stages on HOST buffers (copies inside): the unit-level boundary.
 * ---------------------------------------------------------------------------------------------- */

/* Overlap::trim + Overlap::type (overlap.cpp:117-259) for n independent records: coordinates are
 * trimmed in place, type_out[i] is RALA_K* or RALA_REJECTED. */
int rala_b200_trim_classify(rala_b200_ctx* ctx, rala_ovl_t* ovl, uint64_t n,
                            const rala_pile_t* piles, uint32_t n_piles, uint8_t* type_out);

/* Graph::remove_transitive_edges (graph.cpp:1281-1318) on an arbitrary edge list:
 * marked_out[e] = 1 for every edge the reference removes; *n_pairs = its return value. */
int rala_b200_transitive_reduce(rala_b200_ctx* ctx, uint32_t n_nodes, uint64_t n_edges,
                                const rala_edge_t* edges, uint8_t* marked_out, uint64_t* n_pairs);

/* The duplicate filter of the front end: Graph::initialize's remove_duplicate_overlaps (graph.cpp:273-303) as driven by
 * the grouping loop (:340-361), for n records in file order.  a_id[i] with bit 31 set (RALA_OVL_INVALID) = the record's
 * names did not resolve (the reference holds nullptr: skipped wherever it stands); bit 31 of b_id[i] (orientation) is
 * ignored; length[i] = Overlap::length() as parsed (PAF column 11, MHAP: the longer span).  valid_out[i] = what
 * is_valid_overlap_[i] holds after the pass: 0 for unresolved records, self overlaps and every record of a query group but
 * the last longest one per target.  device_ms (nullable) receives the kernel time.  SURVEY.md 8(f) row 1, first half; the
 * coverage accumulation of the same pass (:305-321, pile.cpp:274-297) stays host code. */
int rala_b200_filter_duplicates(rala_b200_ctx* ctx, const uint32_t* a_id, const uint32_t* b_id, const uint32_t* length,
                                uint64_t n, uint8_t* valid_out, float* device_ms);

/* ------------------------------------------------------------------------------------------------
 * Graph session: the drop-in for Graph::construct's hot loops and Graph::remove_transitive_edges.
 * Call order (what the patched Graph::construct does):
 *   set_piles, set_hills, set_overlaps        inputs produced by the unchanged Graph::initialize
 *   classify                                  graph.cpp:443-518
 *   get_hill_coverage, [host: Pile::break_over_chimeric_hills], set_piles      :704-720
 *   retrim                                    :722-736
 *   loop { get_connections, [host: components, medians, Pile::break_over_chimeric_pits], set_piles,
 *          retrim_promote(&changed) } until !changed                           :738-829
 *   finalize                                  :831-877
 *   get_piles                                 pile liveness for Sequence trimming / node creation (:527-574)
 *   build                                     :552-632 (ids, edge list, adjacency)
 *   get_edges, get_seq_to_node                re-materialise Node/Edge objects on the host
 *   transitive                                :1281-1318
 *   get_marked                                then the unchanged :1320-1334 on the host
 * rala_b200_graph_run() chains classify..transitive with the pile table frozen and no host
 * synchronisation in between (clean data: no hills, no pits).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t n_records;      /* as set */
    uint64_t n_overlaps;     /* current `overlaps` list */
    uint64_t n_internals;    /* current `internals` list */
    uint64_t n_candidates;   /* containment events of the first pass (graph.cpp:469-480) */
    uint32_t n_rounds;       /* fixed-point rounds their resolution took */
    uint32_t n_piles, n_alive_piles;
    uint32_t n_nodes;        /* nodes_.size() */
    uint64_t n_edges;        /* edges_.size() */
    uint64_t n_two_hop;      /* H = sum over edges (a->b) of outdegree(b): two-hop visits of the transitive pass */
    uint64_t n_transitive_pairs; /* return value of remove_transitive_edges */
    uint32_t n_heavy_items;  /* block-per-node work items of the last transitive pass */
    uint32_t n_final_rounds; /* fixed-point rounds of the final containment pass (graph.cpp:831-866) */
    uint64_t n_final_candidates; /* containment events of the final pass */
} rala_b200_counts_t;

int rala_b200_graph_create(rala_b200_ctx* ctx, rala_b200_graph** out);
void rala_b200_graph_destroy(rala_b200_graph* g);

int rala_b200_graph_set_overlaps(rala_b200_graph* g, const rala_ovl_t* ovl, uint64_t n);
/* The same records as six HOST COLUMNS in the layout the kernels read (replaces the same members of rala::Overlap,
 * overlap.hpp:104-116, marshalled column-wise by the host shim): 24 bytes per record cross PCIe instead of 28 and
 * land in place, with no staging copy and no transpose kernel.  Bit 31 of a_id[i] marks an INVALID record
 * (RALA_OVL_INVALID), bit 31 of b_id[i] is the orientation (RALA_OVL_RC); ids are < 2^31.  File order as above. */
int rala_b200_graph_set_overlaps_columns(rala_b200_graph* g, const uint32_t* a_id, const uint32_t* b_id,
                                         const uint32_t* a_begin, const uint32_t* a_end, const uint32_t* b_begin,
                                         const uint32_t* b_end, uint64_t n);
/* The same records in a COMPACT host form for the PCIe link: 12 bytes per record instead of 24 (the upload is 85 % of an
 * end-to-end step).  It relies on two properties of the data the reference consumes: a PAF / MHAP lists the overlaps
 * of one query together (graph.cpp:343-350 assumes the same grouping), and read coordinates of today's long reads fit
 * 16 bits (the caller checks: every coordinate < 65 536, otherwise it uses the column form).
 *   group k = records [group_end[k-1], group_end[k]) (group_end[-1] = 0), all with a_id = query_id[k] (< 2^31)
 *   b_id[i]   as in the column form (bit 31 = orientation)
 *   a_span[i] = a_begin | a_end << 16,   b_span[i] = b_begin | b_end << 16
 * An INVALID record (RALA_OVL_INVALID) is passed with both spans 0: Overlap::trim rejects it in every pass
 * (overlap.cpp:139-142), which is all "invalid" means on this path.  Expanded on the device into the column layout. */
int rala_b200_graph_set_overlaps_packed(rala_b200_graph* g, const uint32_t* query_id, const uint32_t* group_end, uint32_t n_groups,
                                        const uint32_t* b_id, const uint32_t* a_span, const uint32_t* b_span, uint64_t n);
int rala_b200_graph_set_piles(rala_b200_graph* g, const rala_pile_t* piles, const uint8_t* flags /* nullable */,
                              uint32_t n_piles);
/* Optional: have build / transitive / run write their results STRAIGHT into caller memory the GPU can address
 * (cudaHostAlloc / cudaHostRegister'ed host memory, or device memory): the rala_edge_t rows (graph.cpp:576-632)
 * leave as soon as the edge list exists, on a forked stream beside the CSR build and the transitive pass, and the
 * removed-edge marks (graph.cpp:1305-1309) as they are finalised.  Nothing beyond the given capacities (in edges)
 * is written; the counts come from rala_b200_graph_counts.  The buffers hold valid data once the stream has been
 * synchronised (rala_b200_graph_counts / rala_b200_synchronize).  NULL switches an output off again; pageable
 * memory is refused (RALA_B200_ERR_ARG): use get_edges / get_marked for it.  The caller keeps the buffers alive
 * (and pinned) until it has switched them off or destroyed the session: every later build / run writes into them. */
int rala_b200_graph_set_outputs(rala_b200_graph* g, rala_edge_t* edges_out, uint64_t edges_cap, uint8_t* marked_out,
                                uint64_t marked_cap);
int rala_b200_graph_set_hills(rala_b200_graph* g, const rala_hill_t* hills, uint32_t n_hills);

int rala_b200_graph_classify(rala_b200_graph* g);
int rala_b200_graph_retrim(rala_b200_graph* g);
int rala_b200_graph_retrim_promote(rala_b200_graph* g, int* is_changed);
int rala_b200_graph_finalize(rala_b200_graph* g);
int rala_b200_graph_build(rala_b200_graph* g);
int rala_b200_graph_transitive(rala_b200_graph* g);
int rala_b200_graph_run(rala_b200_graph* g);
/* rala_b200_graph_run() replays a captured CUDA graph of the whole chain once it has run a session shape (buffers,
 * sizes, first run after set_piles or not) eagerly: one launch instead of ~45 dependent stream operations.  The
 * per-stage timers of rala_b200_graph_stage_ms() are not recorded inside a graph: pass enabled = 0 before the
 * runs you want stage times for.  Default: enabled. */
int rala_b200_graph_use_cuda_graph(rala_b200_graph* g, int enabled);

/* synchronises the stream and reads the device-side counters */
int rala_b200_graph_counts(rala_b200_graph* g, rala_b200_counts_t* out);

int rala_b200_graph_get_hill_coverage(rala_b200_graph* g, uint32_t* cov_out /* n_hills */);
int rala_b200_graph_get_piles(rala_b200_graph* g, rala_pile_t* piles_out /* n_piles */);
/* (a_id, b_id) of every entry of `overlaps`, for the host's component search (graph.cpp:740-744) */
int rala_b200_graph_get_connections(rala_b200_graph* g, uint32_t* ab_out /* 2 * n_overlaps */);
int rala_b200_graph_get_lists(rala_b200_graph* g, rala_ovl_t* overlaps_out /* nullable */,
                              rala_ovl_t* internals_out /* nullable */);
/* Replaces the device `overlaps` list after finalize by a host-filtered copy of it, same order: what
 * Graph::preprocess(overlaps, sensitive_overlaps_path) (graph.cpp:523, 882-1054, the -s option) leaves. */
int rala_b200_graph_set_kept_overlaps(rala_b200_graph* g, const rala_ovl_t* kept, uint64_t n);
int rala_b200_graph_get_seq_to_node(rala_b200_graph* g, uint32_t* out /* n_piles; 0xFFFFFFFF = dead */);
int rala_b200_graph_get_edges(rala_b200_graph* g, rala_edge_t* out /* n_edges */);
int rala_b200_graph_get_marked(rala_b200_graph* g, uint8_t* out /* n_edges */);
/* Adjacency VIEW, built on the device: off_out[v] .. off_out[v + 1] delimit, in ids_out, the ids of node v's out-edges (which = 0:
 * suffix_edges_, graph.cpp:603 / 625) or in-edges (which = 1: prefix_edges_, :605-606 / 622-624) in ascending edge id, as the
 * reference keeps them; with skip_marked != 0 (after transitive) without the removed edges: the adjacency
 * Graph::remove_marked_objects leaves (graph.cpp:2118-2151).  off_out: n_nodes + 1 words, ids_out: n_edges words. */
int rala_b200_graph_get_adjacency(rala_b200_graph* g, int which, int skip_marked, uint32_t* off_out, uint32_t* ids_out, uint64_t* n_ids_out);

/* Device time of the last run of each stage, CUDA events on the context's stream (ms).
 * Order: classify, retrim, finalize, build, transitive (whole stages), then single kernels:
 * K1 events kernel (first pass over the records), K1b containment fixed point, K3 transitive kernels
 * (light + heavy), K1 survivors kernel (second pass over the records).  For bench.py's roofline object. */
#define RALA_B200_N_STAGES 9
int rala_b200_graph_stage_ms(rala_b200_graph* g, float* ms_out /* RALA_B200_N_STAGES */);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU phases: one process per GPU, records sharded by contiguous FILE RANGE (shard r holds
 * records [t0_r, t0_r + n_r)), pile table replicated.  The reference has no distributed code; this
 * is the partition BASELINE.json's north_star asks for: work split by overlap range (classify) and
 * by source-node range (transitive), the event lists and the edge list all-gathered, the marks
 * merged with an all-reduce(max) (= OR on 0/1 bytes).  The collectives are the caller's
 * (rala_b200/multi.py uses torch.distributed / NCCL); every pointer below is a DEVICE pointer of
 * the caller's exchange buffer, laid out as three columns of `stride` 32-bit words
 * (events: victim | container | time;  edges: src | dst | len).
 *
 *   set_shard; set_piles; set_overlaps(local records)
 *   phase_events              -> events_count, export_events, [all-gather], import_events (all ranks' blocks)
 *   phase_resolve(1)          ordered containment of graph.cpp:469-480 on the union of the events (replicated)
 *   phase_survivors           -> list_counts, [all-gather counts]
 *   phase_final_events(ovl_base, int_base)  time = global position in overlaps ++ internals (graph.cpp:831-866)
 *                             -> events_count, export_events, [all-gather], import_events
 *   phase_resolve(0)
 *   phase_emit_edges          -> export_edges, [all-gather], import_edges (global edge-id order = rank order)
 *   phase_csr                 CSR of the whole graph, replicated (two-hop lookups cross shards)
 *   phase_transitive          T(e) for candidate edges whose source node this rank owns
 *   export_marks, [all-reduce(max)], phase_marks     marked(e) = T(e) | T(e^1)
 * ---------------------------------------------------------------------------------------------- */
/* a context that enqueues on the caller's CUDA stream (cudaStream_t), e.g. torch's current stream */
int rala_b200_create_on_stream(rala_b200_ctx** out, int device, void* cuda_stream);
int rala_b200_graph_set_shard(rala_b200_graph* g, uint32_t t0, int rank, int world);
int rala_b200_graph_phase_events(rala_b200_graph* g);
int rala_b200_graph_events_count(rala_b200_graph* g, uint32_t* n);
int rala_b200_graph_export_events(rala_b200_graph* g, uint32_t* d_cols, uint32_t stride, uint32_t n);
int rala_b200_graph_import_events(rala_b200_graph* g, const uint32_t* d_cols, uint32_t stride, uint32_t n,
                                  uint32_t offset, uint32_t total);
int rala_b200_graph_phase_resolve(rala_b200_graph* g, int first_pass);
int rala_b200_graph_phase_survivors(rala_b200_graph* g);
int rala_b200_graph_list_counts(rala_b200_graph* g, uint32_t* n_ovl, uint32_t* n_int);
int rala_b200_graph_phase_final_events(rala_b200_graph* g, uint32_t ovl_base, uint32_t int_base);
int rala_b200_graph_phase_emit_edges(rala_b200_graph* g, uint32_t* n_local_edges);
int rala_b200_graph_export_edges(rala_b200_graph* g, uint32_t* d_cols, uint32_t stride, uint32_t n);
int rala_b200_graph_import_edges(rala_b200_graph* g, const uint32_t* d_cols, uint32_t stride, uint32_t n,
                                 uint32_t offset, uint32_t total);
int rala_b200_graph_phase_csr(rala_b200_graph* g);
/* Capacity-bounded variants of the exchange steps: element counts travel inside the blocks and the time bases are
 * computed on the device, so no call below synchronises with the host (the sized variants above read counts back).
 *   kind 0, containment events: block (3 * cap + 4 words) = [n clamped to cap | overflow flag | 0 | 0 | victim [cap] | container [cap] | time [cap]]
 *   kind 1, edges: the same layout with src | dst | len
 *   rala_b200_exchange_block_words(kind, cap) is the block size in 32-bit words.  d_gathered = the `world` blocks of an all-gather, in rank order.
 * A count beyond `cap` raises the session's overflow flag (reported by rala_b200_graph_counts): re-run sized.
 * rala_b200_graph_phase_emit_edges accepts n_local_edges == NULL (no read-back). */
uint64_t rala_b200_exchange_block_words(int kind, uint32_t cap);
int rala_b200_graph_export_padded(rala_b200_graph* g, int kind, uint32_t* d_block, uint32_t cap);
int rala_b200_graph_import_gathered(rala_b200_graph* g, int kind, const uint32_t* d_gathered, uint32_t cap, int world);
int rala_b200_graph_export_list_counts(rala_b200_graph* g, uint32_t* d_pair /* n_overlaps, n_internals */);
int rala_b200_graph_phase_final_events_gathered(rala_b200_graph* g, const uint32_t* d_counts /* world x 2 */, int world);
int rala_b200_graph_phase_transitive(rala_b200_graph* g);
int rala_b200_graph_export_marks(rala_b200_graph* g, uint8_t* d_T, uint32_t n);
int rala_b200_graph_phase_marks(rala_b200_graph* g, const uint8_t* d_T, uint32_t n);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU session (ABI version 2): Graph::construct's hot loops + Graph::remove_transitive_edges on `world` ranks,
 * one rank per GPU, with the whole orchestration INSIDE the library.  The reference has no distributed code
 * (rvaser/rala is one process); this is the partition BASELINE.json's north_star names: overlap records by
 * contiguous FILE RANGE (classification), piles / nodes by id range (ordered containment of graph.cpp:469-480 and
 * 831-866, adjacency rows of :576-632, the two-hop test of :1281-1318), CSR of the whole graph replicated.
 *
 * Exchange steps are kernels that write straight into the peer GPUs' memory over NVLink (peer access inside one
 * process, CUDA IPC between processes) separated by device-side flag barriers: one step enqueues kernels only, has
 * no host synchronisation and no library collective, and is replayed as one CUDA graph per rank.
 *
 *   single process, n GPUs (the drop-in CLI, host/graph_b200.cpp):
 *       rala_b200_multi_create(&m, devices, n, 0, n); set_piles; set_overlaps(k, shard k) for every k;
 *       rala_b200_multi_plan(m);  rala_b200_multi_run(m);  get_* ...
 *   one process per GPU (torch.distributed / MPI launchers; rala_b200/multi.py):
 *       rala_b200_multi_create(&m, &device, 1, rank, world); set_piles; set_overlaps(0, own shard);
 *       default_caps -> [max over ranks] -> reserve -> export_handle -> [all-gather handles] -> import_handles;
 *       run; synchronize; demand -> [max over ranks]: re-reserve with larger capacities if something did not fit.
 *
 * A device id may be listed more than once (several ranks share a GPU): that is how the parity tests cover
 * world > 1 on a one-GPU box.  Restrictions: frozen pile table (no hills, no host pile breaking between the
 * passes: the clean-data chain rala_b200_graph_run executes); read ids, coordinates and counts 32-bit as above.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rala_b200_multi rala_b200_multi;

#define RALA_B200_MAX_RANKS 16
/* capacities of the exchange buffers (identical on every rank) */
enum { RALA_B200_CAP_EVENTS = 0,   /* containment events one rank sends to one owner */
       RALA_B200_CAP_EDGES,        /* edges one rank sends to one owner */
       RALA_B200_CAP_SLICE,        /* edges of one rank's CSR slice */
       RALA_B200_CAP_ROUNDS,       /* sweeps after which the resolution of the first containment pass gives up (it ends by itself) */
       RALA_B200_CAP_FINAL_ROUNDS, /* the same for the final containment pass */
       RALA_B200_CAP_LOCAL_EDGES,  /* edges one rank emits */
       RALA_B200_N_CAPS };

typedef struct {
    int world, n_local;
    uint64_t n_records;            /* records of the local ranks */
    uint32_t n_piles, n_alive_piles, n_nodes;
    uint64_t n_edges;              /* edges_.size(): all ranks (every rank knows the total) */
    uint64_t n_local_edges;        /* edges emitted by the local ranks */
    uint64_t n_candidates, n_final_candidates;   /* containment events emitted by the local ranks */
    uint32_t n_rounds, n_final_rounds;           /* resolution rounds needed (same on every rank) */
    uint64_t n_two_hop, n_transitive_pairs;      /* local ranks' share (sum over all ranks = the reference's values) */
    uint32_t n_heavy_items;
    uint32_t fabric_error;         /* 0, or bits: 1 barrier timeout, 2 rounds exhausted, 4 exchange buffer too small */
} rala_b200_multi_counts_t;

/* ranks [first_rank, first_rank + n_local) of `world` live in this process, local rank k on devices[k] */
int rala_b200_multi_create(rala_b200_multi** out, const int* devices, int n_local, int first_rank, int world);
void rala_b200_multi_destroy(rala_b200_multi* m);
const char* rala_b200_multi_last_error(const rala_b200_multi* m);
/* replicated pile table (every local rank gets a copy); as rala_b200_graph_set_piles */
int rala_b200_multi_set_piles(rala_b200_multi* m, const rala_pile_t* piles, const uint8_t* flags, uint32_t n_piles);
/* shard of local rank k: records [t0, t0 + n) of the file, in file order */
int rala_b200_multi_set_overlaps(rala_b200_multi* m, int k, const rala_ovl_t* ovl, uint64_t n, uint64_t t0);
int rala_b200_multi_set_overlaps_columns(rala_b200_multi* m, int k, const uint32_t* a_id, const uint32_t* b_id, const uint32_t* a_begin,
                                         const uint32_t* a_end, const uint32_t* b_begin, const uint32_t* b_end, uint64_t n, uint64_t t0);
int rala_b200_multi_set_overlaps_packed(rala_b200_multi* m, int k, const uint32_t* query_id, const uint32_t* group_end, uint32_t n_groups,
                                        const uint32_t* b_id, const uint32_t* a_span, const uint32_t* b_span, uint64_t n, uint64_t t0);
/* as rala_b200_graph_set_outputs, for the edges local rank k emits (edge ids rala_b200_multi_edge_range) and their marks */
int rala_b200_multi_set_outputs(rala_b200_multi* m, int k, rala_edge_t* edges_out, uint64_t edges_cap, uint8_t* marked_out,
                                uint64_t marked_cap);
/* capacities this process would choose for its own shards (take the maximum over all processes) */
int rala_b200_multi_default_caps(rala_b200_multi* m, uint64_t* caps /* RALA_B200_N_CAPS */);
/* (re)allocate the exchange arenas; with all ranks in one process this also connects them */
int rala_b200_multi_reserve(rala_b200_multi* m, const uint64_t* caps /* RALA_B200_N_CAPS */);
/* change the sweep limits of the two containment resolutions without touching the arenas */
int rala_b200_multi_set_rounds(rala_b200_multi* m, uint32_t rounds, uint32_t final_rounds);
/* one process per GPU: 64-byte CUDA IPC handle of local rank k's arena / the handles of all `world` ranks in rank order */
int rala_b200_multi_export_handle(rala_b200_multi* m, int k, void* handle64);
int rala_b200_multi_import_handles(rala_b200_multi* m, const void* handles /* world x 64 bytes */);
/* one step on every local rank: classify .. transitive.  Enqueues only (replays one CUDA graph per rank from the
 * third call with the same inputs on). */
int rala_b200_multi_run(rala_b200_multi* m);
int rala_b200_multi_use_cuda_graph(rala_b200_multi* m, int enabled);
/* How long a rank waits for its peers at a device-side barrier before it gives up (default 10 000 ms).  After a timeout
 * the step's results are void, no later barrier waits, and demand / counts report RALA_B200_ERR_CUDA. */
int rala_b200_multi_set_barrier_timeout_ms(rala_b200_multi* m, uint32_t ms);
int rala_b200_multi_synchronize(rala_b200_multi* m);
/* after synchronize: what the last step needed (per capacity; rounds: rounds needed) and whether everything fit */
int rala_b200_multi_demand(rala_b200_multi* m, uint64_t* need /* RALA_B200_N_CAPS */, int* fits);
/* all ranks in one process: reserve with the default capacities, run, grow what did not fit, until a step fits */
int rala_b200_multi_plan(rala_b200_multi* m);
int rala_b200_multi_counts(rala_b200_multi* m, rala_b200_multi_counts_t* out);
/* results of local rank k: the edges it emitted are edge ids [*first_edge, *first_edge + *n) (rows in id order) */
int rala_b200_multi_edge_range(rala_b200_multi* m, int k, uint64_t* first_edge, uint64_t* n);
int rala_b200_multi_get_edges(rala_b200_multi* m, int k, rala_edge_t* out);
int rala_b200_multi_get_marked(rala_b200_multi* m, int k, uint8_t* out);
int rala_b200_multi_get_seq_to_node(rala_b200_multi* m, uint32_t* out /* n_piles */);
int rala_b200_multi_get_piles(rala_b200_multi* m, rala_pile_t* out /* n_piles */);
/* CUDA events on every local rank's stream; elapsed = max over the local ranks */
int rala_b200_multi_event_record(rala_b200_multi* m, int which);
int rala_b200_multi_event_elapsed_ms(rala_b200_multi* m, float* ms);
uint64_t rala_b200_multi_launch_count(const rala_b200_multi* m);
/* Diagnostics: device timestamps (ns) of the last barriers of local rank k, oldest first, as pairs (barrier kernel started,
 * every peer had arrived); out holds 2 x 128 values, *n_out = number of valid pairs.  Synchronises. */
int rala_b200_multi_barrier_log(rala_b200_multi* m, int k, uint64_t* out, uint32_t* n_out);
/* Diagnostics: the sweeps of the last containment resolution (pass 0 first, 1 final) of local rank k as pairs (victims still
 * open when the sweep started, ns since the kernel started when it ended); out holds 2 x 48 values. Synchronises. */
int rala_b200_multi_sweep_log(rala_b200_multi* m, int k, int pass, uint64_t* out, uint32_t* n_out);
/* device time of the stages of the last EAGER step of local rank k (as rala_b200_graph_stage_ms) */
int rala_b200_multi_stage_ms(rala_b200_multi* m, int k, float* ms_out /* RALA_B200_N_STAGES */);

#ifdef __cplusplus
}
#endif
#endif
