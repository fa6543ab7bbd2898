// tests/host_dupfilter.cu — CPU check of rala_b200/csrc/common.cuh's duplicate_filter_keeps() (what every thread of
// k_filter_duplicates evaluates on the GPU: "the last longest record per (query group, target) survives") against the
// ORACLE's literal restatement of the reference's nested loops (oracle/rala_oracle.c ora_filter_duplicates, pinned to
// the compiled reference by tests/golden/dups.npz).  Host-only program: nvcc compiles the __host__ __device__ function
// for the CPU; no GPU is touched.  Built and run by tests/test_frontend.py.
//
//   host_dupfilter <sequences> <seed>  ->  one JSON line {"sequences": .., "records": .., "kept": .., "mismatches": ..}
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../oracle/rala_oracle.h"
#include "../rala_b200/csrc/common.cuh"

static uint64_t rng_state;
static inline uint64_t rnd() {   // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t below(uint32_t n) { return n ? (uint32_t) (rnd() % n) : 0u; }

int main(int argc, char** argv) {
    const uint64_t sequences = argc > 1 ? strtoull(argv[1], nullptr, 10) : 20000ull;
    rng_state = argc > 2 ? strtoull(argv[2], nullptr, 10) : 1ull;
    uint64_t records = 0, kept = 0, mismatches = 0;
    std::vector<uint32_t> a, b, len;
    std::vector<uint8_t> want;
    for (uint64_t it = 0; it < sequences; ++it) {
        // a short file: a handful of query groups drawn from tiny id / length pools, so that duplicates, ties, self
        // overlaps, a query coming back later and unresolved records (also at the very start / end, also whole groups)
        // all happen all the time
        const uint32_t n = 1u + below(60u), ids = 2u + below(5u), lens = 1u + below(4u), ghost_pct = below(4) == 0 ? 30u : 5u;
        a.assign(n, 0); b.assign(n, 0); len.assign(n, 0); want.assign(n, 0);
        uint32_t cur = below(ids);
        for (uint32_t i = 0; i < n; ++i) {
            if (below(6) == 0) cur = below(ids);                 // next group (possibly the same query again: still one group)
            a[i] = below(100) < ghost_pct ? (0x80000000u | below(ids)) : cur;
            b[i] = below(ids) | (below(2) << 31);                // orientation bit must be ignored
            len[i] = 1000u + below(lens);
        }
        ora_filter_duplicates(a.data(), b.data(), len.data(), n, want.data());
        for (uint32_t k = 0; k < n; ++k) {
            const bool got = rb::duplicate_filter_keeps(a.data(), b.data(), len.data(), n, k);
            mismatches += got != (want[k] != 0);
            kept += got;
        }
        records += n;
    }
    printf("{\"sequences\": %llu, \"records\": %llu, \"kept\": %llu, \"mismatches\": %llu}\n", (unsigned long long) sequences,
           (unsigned long long) records, (unsigned long long) kept, (unsigned long long) mismatches);
    return mismatches ? 1 : 0;
}
