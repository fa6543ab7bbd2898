// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Builds into oracle/_ref/rala_ref together with the UNMODIFIED reference
// sources where they lie under /root/reference (see oracle/Makefile). Nothing
// of the reference is copied into this repository: its graph.cpp is pulled in
// with an #include (after `#define private public`, so the private state that
// parity is defined on — piles_, nodes_, edges_, is_valid_overlap_ — can be
// read and, for the `hotpath` mode, injected).
//
// Modes
//   dump    <reads.fa> <ovl.paf> <out_prefix> [threads]
//       Ground truth. Runs the reference's own Graph::initialize() on one Graph
//       (to export the hot path's INPUTS: parsed overlap records, validity
//       mask, pile table, flags, hills), the reference's own
//       Graph::construct("") on a second Graph (to export the edge list) and
//       its own Graph::remove_transitive_edges() (removed set, count,
//       transitive_edges_).  Then re-runs the hot path with this file's
//       staged driver (below) on a third Graph, dumping the state at every
//       stage boundary, and FAILS (exit 3) unless the staged driver's edge list
//       and removed set equal the reference's own bit for bit.
//   hotpath <in_prefix> <out_prefix> [repeat]
//       Binary inputs (records, piles, hills) -> reference objects built
//       through the reference's private constructors -> the same staged driver
//       with the pile table frozen (no Pile::break_over_*; needs no per-base
//       coverage vectors) -> the reference's own remove_transitive_edges().
//       Prints one JSON line with per-phase wall times: this is the CPU
//       baseline (`bench.py --impl reference`, cpu_baseline.kind="reference").
//
// The staged driver is the reference's control flow for graph.cpp:443-518,
// 699-880 and 552-632 re-stated around the reference's OWN Overlap::trim,
// Overlap::type, Pile::check_chimeric_hills, Pile::break_over_*,
// Graph::Node/Edge and Graph::remove_transitive_edges — the loops had to be
// re-stated because Graph::construct is one monolithic function that parses
// text inside the loop.  `dump` proves the re-statement on every input it sees.

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#define private public
#include "graph.cpp"  // -I/root/reference/src ; brings Graph::Node / Graph::Edge into scope
#undef private

namespace {

using rala::Graph;
using rala::Overlap;
using rala::OverlapType;
using rala::Pile;
using OvlVec = std::vector<std::unique_ptr<Overlap>>;
using Clock = std::chrono::steady_clock;

double seconds_since(Clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::duration<double>>(Clock::now() - t0).count();
}

void write_u32(const std::string& path, const std::vector<uint32_t>& v) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "[rala_ref] cannot write %s\n", path.c_str()); exit(2); }
    if (!v.empty()) fwrite(v.data(), sizeof(uint32_t), v.size(), f);
    fclose(f);
}

std::vector<uint32_t> read_u32(const std::string& path, bool optional = false) {
    std::vector<uint32_t> v;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        if (optional) return v;
        fprintf(stderr, "[rala_ref] cannot read %s\n", path.c_str()); exit(2);
    }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize(bytes / sizeof(uint32_t));
    if (!v.empty() && fread(v.data(), sizeof(uint32_t), v.size(), f) != v.size()) {
        fprintf(stderr, "[rala_ref] short read %s\n", path.c_str()); exit(2);
    }
    fclose(f);
    return v;
}

// rala_ovl_t layout (include/rala_b200.h): a_id b_id a_begin a_end b_begin b_end flags
void append_record(std::vector<uint32_t>& dst, const Overlap& o, uint32_t extra_flags = 0) {
    dst.push_back(static_cast<uint32_t>(o.a_id_));
    dst.push_back(static_cast<uint32_t>(o.b_id_));
    dst.push_back(o.a_begin_);
    dst.push_back(o.a_end_);
    dst.push_back(o.b_begin_);
    dst.push_back(o.b_end_);
    dst.push_back((o.orientation_ & 1u) | extra_flags);
}

void dump_list(const std::string& path, const OvlVec& list) {
    std::vector<uint32_t> out;
    out.reserve(list.size() * 7);
    for (const auto& it : list) {
        if (it != nullptr) append_record(out, *it);
    }
    write_u32(path, out);
}

// n x 4: begin end flags(bit0 has_chimeric_hill, bit1 has_chimeric_region) median ; dead pile = 0 0 0 0
void dump_piles(const std::string& path, const std::vector<std::unique_ptr<Pile>>& piles) {
    std::vector<uint32_t> out;
    out.reserve(piles.size() * 4);
    for (const auto& p : piles) {
        if (p == nullptr) {
            out.insert(out.end(), {0u, 0u, 0u, 0u});
        } else {
            uint32_t flags = (p->has_chimeric_hill() ? 1u : 0u) | (p->has_chimeric_region() ? 2u : 0u);
            out.insert(out.end(), {p->begin(), p->end(), flags, static_cast<uint32_t>(p->median())});
        }
    }
    write_u32(path, out);
}

// rows: pile first second coverage   (ascending pile id, hills in the pile's own order)
void dump_hills(const std::string& path, const std::vector<std::unique_ptr<Pile>>& piles) {
    std::vector<uint32_t> out;
    for (const auto& p : piles) {
        if (p == nullptr) continue;
        for (size_t i = 0; i < p->chimeric_hills_.size(); ++i) {
            out.insert(out.end(), {static_cast<uint32_t>(p->id()), p->chimeric_hills_[i].first,
                p->chimeric_hills_[i].second, p->chimeric_hill_coverage_[i]});
        }
    }
    write_u32(path, out);
}

void dump_edges(const std::string& prefix, Graph& g) {
    std::vector<uint32_t> edges, node_seq;
    edges.reserve(g.edges_.size() * 3);
    for (const auto& e : g.edges_) {
        edges.insert(edges.end(), {static_cast<uint32_t>(e->begin_node_->id_),
            static_cast<uint32_t>(e->end_node_->id_), e->length_});
    }
    for (const auto& n : g.nodes_) {
        node_seq.push_back(static_cast<uint32_t>(n->sequence_ids_[0]));
    }
    write_u32(prefix + ".edges.u32", edges);
    write_u32(prefix + ".node_seq.u32", node_seq);
}

// adjacency as the reference holds it: per node, suffix edge ids then prefix edge ids
void dump_adjacency(const std::string& prefix, Graph& g) {
    std::vector<uint32_t> off, ids;
    off.push_back(0);
    for (const auto& n : g.nodes_) {
        if (n != nullptr) {
            for (const auto& e : n->suffix_edges_) ids.push_back(static_cast<uint32_t>(e->id_));
        }
        off.push_back(static_cast<uint32_t>(ids.size()));
    }
    write_u32(prefix + ".suffix_off.u32", off);
    write_u32(prefix + ".suffix_ids.u32", ids);
    off.assign(1, 0u);
    ids.clear();
    for (const auto& n : g.nodes_) {
        if (n != nullptr) {
            for (const auto& e : n->prefix_edges_) ids.push_back(static_cast<uint32_t>(e->id_));
        }
        off.push_back(static_cast<uint32_t>(ids.size()));
    }
    write_u32(prefix + ".prefix_off.u32", off);
    write_u32(prefix + ".prefix_ids.u32", ids);
}

uint32_t reduce_and_dump(const std::string& prefix, Graph& g) {
    uint32_t n_pairs = g.remove_transitive_edges();  // the reference's own function
    std::vector<uint32_t> removed(g.edges_.size());
    for (size_t i = 0; i < g.edges_.size(); ++i) removed[i] = g.edges_[i] == nullptr ? 1u : 0u;
    write_u32(prefix + ".removed.u32", removed);
    std::vector<uint32_t> te;
    for (const auto& it : g.transitive_edges_) {
        te.push_back(static_cast<uint32_t>(it.first));
        te.push_back(static_cast<uint32_t>(it.second));
    }
    write_u32(prefix + ".transitive_pairs.u32", te);
    return n_pairs;
}

// ---------------------------------------------------------------------------
// Staged driver: control flow of graph.cpp:443-518 / 699-880 / 552-632 around
// the reference's own member functions.
// ---------------------------------------------------------------------------
struct StagedDriver {
    Graph& g;
    bool break_piles;          // false => pile table frozen (hotpath mode)
    std::string stage_prefix;  // non-empty => dump state at stage boundaries
    OvlVec overlaps, internals;
    uint32_t pit_rounds = 0;
    double t_classify = 0, t_preprocess = 0, t_nodes = 0, t_edges = 0;

    StagedDriver(Graph& graph, bool brk, const std::string& sp)
        : g(graph), break_piles(brk), stage_prefix(sp) {}

    void stage_dump(const std::string& tag) {
        if (stage_prefix.empty()) return;
        dump_list(stage_prefix + "." + tag + ".ovl.u32", overlaps);
        dump_list(stage_prefix + "." + tag + ".int.u32", internals);
        dump_piles(stage_prefix + "." + tag + ".piles.u32", g.piles_);
    }

    // graph.cpp:448-517.  `all` holds every record in file order; `gate(i, o)`
    // is the is_valid_overlap_/transmute part of the condition at :450-451.
    void classify(OvlVec& all, const std::function<bool(uint64_t, Overlap&)>& gate) {
        auto t0 = Clock::now();
        auto& piles = g.piles_;
        for (uint64_t i = 0; i < all.size(); ++i) {
            auto& it = all[i];
            if (!gate(i, *it) || !it->trim(piles)) {
                it.reset();
                continue;
            }
            if (piles[it->a_id()]->has_chimeric_hill()) piles[it->a_id()]->check_chimeric_hills(it);
            if (piles[it->b_id()]->has_chimeric_hill()) piles[it->b_id()]->check_chimeric_hills(it);

            switch (it->type(piles)) {
                case OverlapType::kX:
                    internals.emplace_back(std::move(it));
                    break;
                case OverlapType::kB:
                    if (!piles[it->b_id()]->has_chimeric_region()) {
                        piles[it->a_id()].reset();
                        it.reset();
                    }
                    break;
                case OverlapType::kA:
                    if (!piles[it->a_id()]->has_chimeric_region()) {
                        piles[it->b_id()].reset();
                        it.reset();
                    }
                    break;
                default:
                    break;
            }
        }
        for (auto& it : all) {
            if (it != nullptr && piles[it->a_id()] != nullptr && piles[it->b_id()] != nullptr) {
                overlaps.emplace_back(std::move(it));
            }
        }
        all.clear();
        rala::shrinkToFit(overlaps, 0);
        for (auto& it : internals) {
            if (piles[it->a_id()] == nullptr || piles[it->b_id()] == nullptr) it.reset();
        }
        rala::shrinkToFit(internals, 0);
        t_classify += seconds_since(t0);
        if (!stage_prefix.empty()) {
            stage_dump("s1");
            dump_hills(stage_prefix + ".s1.hills.u32", piles);
        }
    }

    void retrim(OvlVec& list, bool* changed) {
        for (auto& it : list) {
            if (!it->trim(g.piles_)) {
                it.reset();
                if (changed) *changed = true;
            }
        }
        rala::shrinkToFit(list, 0);
    }

    // connected components over `overlaps` + per-component median (graph.cpp:740-783);
    // any correct component labelling yields the same medians.
    void break_pits() {
        auto& piles = g.piles_;
        std::vector<uint32_t> parent(piles.size());
        std::iota(parent.begin(), parent.end(), 0u);
        std::function<uint32_t(uint32_t)> find = [&](uint32_t x) {
            while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
            return x;
        };
        std::vector<bool> touched(piles.size(), false);
        for (const auto& it : overlaps) {
            uint32_t a = find(it->a_id()), b = find(it->b_id());
            touched[it->a_id()] = touched[it->b_id()] = true;
            if (a != b) parent[std::max(a, b)] = std::min(a, b);
        }
        std::unordered_map<uint32_t, std::vector<uint32_t>> comps;
        for (uint32_t i = 0; i < piles.size(); ++i) {
            if (touched[i]) comps[find(i)].push_back(i);
        }
        for (auto& kv : comps) {
            std::vector<uint16_t> medians;
            for (auto i : kv.second) medians.push_back(piles[i]->median());
            std::nth_element(medians.begin(), medians.begin() + medians.size() / 2, medians.end());
            uint16_t component_median = medians[medians.size() / 2];
            for (auto i : kv.second) {
                if (!piles[i]->break_over_chimeric_pits(component_median)) piles[i].reset();
            }
        }
    }

    // graph.cpp:699-880
    void preprocess() {
        auto t0 = Clock::now();
        auto& piles = g.piles_;
        if (break_piles) {
            for (auto& p : piles) {
                if (p != nullptr && p->has_chimeric_hill() && !p->break_over_chimeric_hills()) p.reset();
            }
        }
        retrim(overlaps, nullptr);
        retrim(internals, nullptr);
        stage_dump("s2");

        while (true) {
            if (break_piles) break_pits();
            bool is_changed = false;
            retrim(overlaps, &is_changed);
            for (auto& it : internals) {
                if (!it->trim(piles)) {
                    it.reset();
                    continue;
                }
                auto t = it->type(piles);
                if (t == OverlapType::kAB || t == OverlapType::kBA) overlaps.emplace_back(std::move(it));
            }
            rala::shrinkToFit(internals, 0);
            stage_dump("s3r" + std::to_string(pit_rounds));
            ++pit_rounds;
            if (!is_changed) break;
        }

        for (auto* list : {&overlaps, &internals}) {
            for (auto& it : *list) {
                if (piles[it->a_id()] == nullptr || piles[it->b_id()] == nullptr) {
                    it.reset();
                    continue;
                }
                auto t = it->type(piles);
                if (t == OverlapType::kA) {
                    piles[it->b_id()].reset();
                    it.reset();
                } else if (t == OverlapType::kB) {
                    piles[it->a_id()].reset();
                    it.reset();
                }
            }
        }
        rala::shrinkToFit(internals, 0);
        for (auto& it : overlaps) {
            if (it != nullptr && (piles[it->a_id()] == nullptr || piles[it->b_id()] == nullptr)) it.reset();
        }
        rala::shrinkToFit(overlaps, 0);
        t_preprocess += seconds_since(t0);
        stage_dump("s4");
    }

    // graph.cpp:552-632 with empty node strings (the bases never enter the path)
    void build_graph() {
        auto t0 = Clock::now();
        auto& piles = g.piles_;
        std::vector<int64_t> s2n(piles.size(), -1);
        uint64_t node_id = 0;
        const std::string empty;
        for (uint64_t i = 0; i < piles.size(); ++i) {
            if (piles[i] == nullptr) continue;
            s2n[i] = node_id;
            std::unique_ptr<Graph::Node> n(new Graph::Node(node_id++, i, empty, empty));
            std::unique_ptr<Graph::Node> nc(new Graph::Node(node_id++, i, empty, empty));
            n->pair_ = nc.get();
            nc->pair_ = n.get();
            g.nodes_.emplace_back(std::move(n));
            g.nodes_.emplace_back(std::move(nc));
        }
        t_nodes += seconds_since(t0);
        t0 = Clock::now();
        uint64_t edge_id = 0;
        for (auto& it : overlaps) {
            auto* na = g.nodes_[s2n[it->a_id()]].get();
            auto* nb = g.nodes_[s2n[it->b_id()] + it->orientation()].get();
            const auto& pa = piles[it->a_id()];
            const auto& pb = piles[it->b_id()];
            uint32_t al = pa->end() - pa->begin(), a0 = it->a_begin() - pa->begin(), a1 = it->a_end() - pa->begin();
            uint32_t bl = pb->end() - pb->begin();
            uint32_t b0 = it->orientation() == 0 ? it->b_begin() - pb->begin() : bl - it->b_end() + pb->begin();
            uint32_t b1 = it->orientation() == 0 ? it->b_end() - pb->begin() : bl - it->b_begin() + pb->begin();
            auto t = it->type(piles);
            Graph::Node *from = nullptr, *to = nullptr;
            uint32_t len = 0, len_c = 0;
            if (t == OverlapType::kAB) {
                from = na; to = nb; len = a0 - b0; len_c = (bl - b1) - (al - a1);
            } else if (t == OverlapType::kBA) {
                from = nb; to = na; len = b0 - a0; len_c = (al - a1) - (bl - b1);
            }
            if (from != nullptr) {
                std::unique_ptr<Graph::Edge> e(new Graph::Edge(edge_id++, from, to, len));
                std::unique_ptr<Graph::Edge> ec(new Graph::Edge(edge_id++, to->pair_, from->pair_, len_c));
                e->pair_ = ec.get();
                ec->pair_ = e.get();
                from->suffix_edges_.emplace_back(e.get());
                from->pair_->prefix_edges_.emplace_back(ec.get());
                to->prefix_edges_.emplace_back(e.get());
                to->pair_->suffix_edges_.emplace_back(ec.get());
                g.edges_.emplace_back(std::move(e));
                g.edges_.emplace_back(std::move(ec));
            }
            it.reset();
        }
        t_edges += seconds_since(t0);
    }
};

std::unique_ptr<Graph> make_graph(const std::string& reads, const std::string& ovl, uint32_t threads) {
    return rala::createGraph(reads, ovl, threads);
}

std::unique_ptr<Graph> make_bare_graph(uint32_t threads) {
    return std::unique_ptr<Graph>(new Graph(nullptr, nullptr, threads));
}

bool same_file_u32(const std::string& a, const std::string& b) {
    return read_u32(a) == read_u32(b);
}

int mode_dump(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: rala_ref dump <reads> <ovl> <out_prefix> [threads]\n"); return 2; }
    std::string reads = argv[2], ovl = argv[3], prefix = argv[4];
    uint32_t threads = argc > 5 ? atoi(argv[5]) : 1;

    // (1) hot-path inputs from the reference's own front end
    uint64_t n_records = 0;
    {
        auto gi = make_graph(reads, ovl, threads);
        gi->initialize();
        dump_piles(prefix + ".in.piles.u32", gi->piles_);
        dump_hills(prefix + ".in.hills.u32", gi->piles_);
        std::vector<uint32_t> lens;
        {
            // read lengths: pile data vectors are sized to the read (pile.cpp:60); dead piles lost theirs,
            // so re-read them from the sequence file
            std::vector<std::unique_ptr<rala::Sequence>> seqs;
            gi->sparser_->reset();
            gi->sparser_->parse_objects(seqs, -1);
            for (const auto& s : seqs) lens.push_back(static_cast<uint32_t>(s->data().size()));
        }
        write_u32(prefix + ".in.read_len.u32", lens);

        OvlVec all;
        gi->oparser_->reset();
        gi->oparser_->parse_objects(all, -1);
        std::vector<uint32_t> rec;
        rec.reserve(all.size() * 7);
        for (uint64_t i = 0; i < all.size(); ++i) {
            auto& o = *all[i];
            bool known = true;
            if (!o.a_name_.empty()) {
                auto f = gi->name_to_id_.find(o.a_name_);
                if (f == gi->name_to_id_.end()) known = false; else o.a_id_ = f->second;
            }
            if (!o.b_name_.empty()) {
                auto f = gi->name_to_id_.find(o.b_name_);
                if (f == gi->name_to_id_.end()) known = false; else o.b_id_ = f->second;
            }
            bool valid = known && gi->is_valid_overlap_[i];
            if (!known) { o.a_id_ = 0xFFFFFFFFu; o.b_id_ = 0xFFFFFFFFu; }
            append_record(rec, o, valid ? 0u : 2u);  // flags bit1 = invalid (is_valid_overlap_ false or unknown name)
        }
        n_records = all.size();
        write_u32(prefix + ".in.records.u32", rec);
    }

    // (2) the reference's own construct + remove_transitive_edges
    uint32_t ref_pairs = 0;
    size_t ref_nodes = 0, ref_edges = 0;
    {
        auto gc = make_graph(reads, ovl, threads);
        gc->construct("");
        dump_edges(prefix + ".ref", *gc);
        dump_adjacency(prefix + ".ref", *gc);
        dump_piles(prefix + ".ref.piles.u32", gc->piles_);
        ref_nodes = gc->nodes_.size();
        ref_edges = gc->edges_.size();
        ref_pairs = reduce_and_dump(prefix + ".ref", *gc);
        dump_adjacency(prefix + ".ref.after", *gc);
    }

    // (3) staged driver on real piles, stage dumps, self-check against (2)
    uint32_t st_pairs = 0;
    uint32_t pit_rounds = 0;
    {
        auto gs = make_graph(reads, ovl, threads);
        gs->initialize();
        OvlVec all;
        gs->oparser_->reset();
        gs->oparser_->parse_objects(all, -1);
        StagedDriver drv(*gs, true, prefix + ".stage");
        drv.classify(all, [&](uint64_t i, Overlap& o) {
            return gs->is_valid_overlap_[i] && o.transmute(gs->piles_, gs->name_to_id_);
        });
        drv.preprocess();
        drv.build_graph();
        pit_rounds = drv.pit_rounds;
        dump_edges(prefix + ".stage", *gs);
        st_pairs = reduce_and_dump(prefix + ".stage", *gs);
    }
    bool ok = same_file_u32(prefix + ".ref.edges.u32", prefix + ".stage.edges.u32") &&
        same_file_u32(prefix + ".ref.removed.u32", prefix + ".stage.removed.u32") &&
        same_file_u32(prefix + ".ref.node_seq.u32", prefix + ".stage.node_seq.u32") &&
        same_file_u32(prefix + ".ref.piles.u32", prefix + ".stage.s4.piles.u32") &&
        ref_pairs == st_pairs;
    printf("{\"mode\": \"dump\", \"records\": %lu, \"nodes\": %zu, \"edges\": %zu, \"transitive_pairs\": %u, "
        "\"pit_rounds\": %u, \"staged_driver_matches_reference\": %s}\n",
        n_records, ref_nodes, ref_edges, ref_pairs, pit_rounds, ok ? "true" : "false");
    return ok ? 0 : 3;
}

int mode_hotpath(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: rala_ref hotpath <in_prefix> <out_prefix|-> [repeat]\n"); return 2; }
    std::string in = argv[2], out = argv[3];
    int repeat = argc > 4 ? atoi(argv[4]) : 1;
    bool write_out = out != "-";

    auto rec = read_u32(in + ".in.records.u32");
    auto pil = read_u32(in + ".in.piles.u32");
    auto hil = read_u32(in + ".in.hills.u32", true);
    auto len = read_u32(in + ".in.read_len.u32");
    uint64_t n = rec.size() / 7, n_piles = pil.size() / 4;

    double best[6] = {1e30, 1e30, 1e30, 1e30, 1e30, 1e30};
    size_t n_nodes = 0, n_edges = 0;
    uint32_t n_pairs = 0;
    for (int r = 0; r < repeat; ++r) {
        auto g = make_bare_graph(1);
        for (uint64_t i = 0; i < n_piles; ++i) {
            if (pil[4 * i + 1] == 0) { g->piles_.emplace_back(nullptr); continue; }
            auto p = rala::createPile(i, 0);  // no per-base coverage vector: the hot path never reads it
            p->begin_ = pil[4 * i];
            p->end_ = pil[4 * i + 1];
            p->median_ = static_cast<uint16_t>(pil[4 * i + 3]);
            if ((pil[4 * i + 2] & 2u) && !(pil[4 * i + 2] & 1u)) p->chimeric_pits_.emplace_back(0u, 0u);  // has_chimeric_region() without hills
            g->piles_.emplace_back(std::move(p));
        }
        for (uint64_t h = 0; h + 3 < hil.size(); h += 4) {
            auto& p = g->piles_[hil[h]];
            if (p == nullptr) continue;
            p->chimeric_hills_.emplace_back(hil[h + 1], hil[h + 2]);
            p->chimeric_hill_coverage_.emplace_back(0u);
        }
        OvlVec all;
        all.reserve(n);
        for (uint64_t i = 0; i < n; ++i) {
            const uint32_t* q = &rec[7 * i];
            uint32_t a = q[0], b = q[1];
            uint32_t al = a < len.size() ? len[a] : 0, bl = b < len.size() ? len[b] : 0;
            // MHAP constructor (overlap.cpp:12-20): ids are 1-based there, orientation = a_rc != b_rc
            std::unique_ptr<Overlap> o(new Overlap(static_cast<uint64_t>(a) + 1, static_cast<uint64_t>(b) + 1, 0.0, 0,
                0, q[2], q[3], al, q[6] & 1u, q[4], q[5], bl));
            o->is_transmuted_ = true;  // ids are already numeric; trim() itself rejects dead piles (overlap.cpp:123-126)
            all.emplace_back(std::move(o));
        }
        StagedDriver drv(*g, false, (write_out && r == 0) ? out + ".stage" : std::string());
        auto t_all = Clock::now();
        drv.classify(all, [&](uint64_t i, Overlap& o) {
            return !(rec[7 * i + 6] & 2u) && o.a_id_ < n_piles && o.b_id_ < n_piles;
        });
        drv.preprocess();
        drv.build_graph();
        if (write_out && r == 0) {
            dump_edges(out + ".stage", *g);
            dump_adjacency(out + ".stage", *g);
        }
        n_nodes = g->nodes_.size();
        n_edges = g->edges_.size();
        auto t0 = Clock::now();
        if (write_out && r == 0) {
            n_pairs = reduce_and_dump(out + ".stage", *g);
        } else {
            n_pairs = g->remove_transitive_edges();
        }
        double t_tr = seconds_since(t0), t_total = seconds_since(t_all);
        double cur[6] = {drv.t_classify, drv.t_preprocess, drv.t_nodes, drv.t_edges, t_tr, t_total};
        for (int k = 0; k < 6; ++k) best[k] = std::min(best[k], cur[k]);
    }
    printf("{\"mode\": \"hotpath\", \"records\": %lu, \"piles\": %lu, \"nodes\": %zu, \"edges\": %zu, \"transitive_pairs\": %u, "
        "\"t_classify\": %.6f, \"t_preprocess\": %.6f, \"t_nodes\": %.6f, \"t_edges\": %.6f, \"t_transitive\": %.6f, "
        "\"t_total\": %.6f, \"repeat\": %d}\n",
        n, n_piles, n_nodes, n_edges, n_pairs, best[0], best[1], best[2], best[3], best[4], best[5], repeat);
    return 0;
}

// unit-level probes of the reference's own pure functions, for the oracle's known-answer tests:
//   trimtype: stdin rows "ab ae bb be ori pa0 pa1 pb0 pb1" -> "ok ab' ae' bb' be' type"
//   comparable: stdin rows "a b" (u32, u32) -> "0|1"  (comparable((double)a,(double)b,0.12), graph.cpp:26-29)
int mode_trimtype() {
    uint32_t v[9];
    while (scanf("%u %u %u %u %u %u %u %u %u", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7], &v[8]) == 9) {
        std::vector<std::unique_ptr<Pile>> piles;
        piles.emplace_back(rala::createPile(0, 0));
        piles.emplace_back(rala::createPile(1, 0));
        piles[0]->begin_ = v[5]; piles[0]->end_ = v[6];
        piles[1]->begin_ = v[7]; piles[1]->end_ = v[8];
        Overlap o(1, 2, 0.0, 0, 0, v[0], v[1], 0xFFFFFFFFu, v[4] & 1u, v[2], v[3], 0xFFFFFFFFu);
        o.is_transmuted_ = true;
        bool ok = o.trim(piles);
        int type = -1;
        if (ok) type = static_cast<int>(o.type(piles));
        printf("%d %u %u %u %u %d\n", ok ? 1 : 0, o.a_begin_, o.a_end_, o.b_begin_, o.b_end_, type);
    }
    return 0;
}

int mode_comparable() {
    uint32_t a, b;
    while (scanf("%u %u", &a, &b) == 2) {
        printf("%d\n", rala::comparable(static_cast<double>(a), static_cast<double>(b), 0.12) ? 1 : 0);
    }
    return 0;
}

// transitive: arbitrary injected graph.  in: <prefix>.edges.u32 (src dst len per edge id, pair(e)=e^1), n_nodes arg.
int mode_transitive(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: rala_ref transitive <edges.u32> <n_nodes> <out_prefix>\n"); return 2; }
    auto e = read_u32(argv[2]);
    uint64_t n_nodes = strtoull(argv[3], nullptr, 10);
    std::string out = argv[4];
    auto g = make_bare_graph(1);
    const std::string empty;
    for (uint64_t i = 0; i < n_nodes; ++i) {
        g->nodes_.emplace_back(new Graph::Node(i, i >> 1, empty, empty));
    }
    for (uint64_t i = 0; i + 1 < n_nodes; i += 2) {
        g->nodes_[i]->pair_ = g->nodes_[i + 1].get();
        g->nodes_[i + 1]->pair_ = g->nodes_[i].get();
    }
    uint64_t n_edges = e.size() / 3;
    for (uint64_t i = 0; i < n_edges; ++i) {
        auto* from = g->nodes_[e[3 * i]].get();
        auto* to = g->nodes_[e[3 * i + 1]].get();
        g->edges_.emplace_back(new Graph::Edge(i, from, to, e[3 * i + 2]));
        from->suffix_edges_.emplace_back(g->edges_.back().get());
        to->prefix_edges_.emplace_back(g->edges_.back().get());
    }
    for (uint64_t i = 0; i + 1 < n_edges; i += 2) {
        g->edges_[i]->pair_ = g->edges_[i + 1].get();
        g->edges_[i + 1]->pair_ = g->edges_[i].get();
    }
    auto t0 = Clock::now();
    uint32_t n_pairs = reduce_and_dump(out, *g);
    printf("{\"mode\": \"transitive\", \"nodes\": %lu, \"edges\": %lu, \"transitive_pairs\": %u, \"t_transitive\": %.6f}\n",
        n_nodes, n_edges, n_pairs, seconds_since(t0));
    return 0;
}

// The duplicate filter of the reference's front end (Graph::initialize, graph.cpp:273-303, run by :328-371): run the
// reference's own initialize() on <reads> + <ovl>, then write one row per overlap record of the file, in file order:
//   a_id b_id length known valid      (u32 each; ids 0xFFFFFFFF and known = 0 when a name is not in <reads>)
// `length` is Overlap::length() as parsed (PAF column 11 / MHAP max span), `valid` is is_valid_overlap_[i].
int mode_dupfilter(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: rala_ref dupfilter <reads> <ovl> <out.u32> [threads]\n"); return 2; }
    std::string reads = argv[2], ovl = argv[3], out = argv[4];
    uint32_t threads = argc > 5 ? atoi(argv[5]) : 1;
    auto gi = make_graph(reads, ovl, threads);
    auto t0 = Clock::now();
    gi->initialize();
    double t_init = seconds_since(t0);
    OvlVec all;
    gi->oparser_->reset();
    gi->oparser_->parse_objects(all, -1);
    std::vector<uint32_t> rows;
    rows.reserve(all.size() * 5);
    uint64_t n_valid = 0;
    for (uint64_t i = 0; i < all.size(); ++i) {
        auto& o = *all[i];
        bool known = true;
        uint32_t a = 0xFFFFFFFFu, b = 0xFFFFFFFFu;
        auto fa = gi->name_to_id_.find(o.a_name_);
        auto fb = gi->name_to_id_.find(o.b_name_);
        if (fa == gi->name_to_id_.end() || fb == gi->name_to_id_.end()) known = false;
        if (known) { a = static_cast<uint32_t>(fa->second); b = static_cast<uint32_t>(fb->second); }
        const bool valid = gi->is_valid_overlap_[i];
        n_valid += valid;
        rows.push_back(a); rows.push_back(b); rows.push_back(o.length()); rows.push_back(known ? 1u : 0u); rows.push_back(valid ? 1u : 0u);
    }
    write_u32(out, rows);
    printf("{\"mode\": \"dupfilter\", \"records\": %lu, \"valid\": %lu, \"t_initialize\": %.6f}\n", all.size(), n_valid, t_init);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: rala_ref dump|hotpath|transitive|trimtype|comparable|dupfilter ...\n");
        return 2;
    }
    std::string mode = argv[1];
    if (mode == "dump") return mode_dump(argc, argv);
    if (mode == "hotpath") return mode_hotpath(argc, argv);
    if (mode == "transitive") return mode_transitive(argc, argv);
    if (mode == "trimtype") return mode_trimtype();
    if (mode == "comparable") return mode_comparable();
    if (mode == "dupfilter") return mode_dupfilter(argc, argv);
    fprintf(stderr, "[rala_ref] unknown mode %s\n", mode.c_str());
    return 2;
}
