// TEST INFRASTRUCTURE.  The reference's OWN CrossCat::get_sorted_groupids (src/cross_cat.cc:212-236, compiled unmodified
// from /root/reference by `make -C oracle ref`) on kinds whose clustering counts are read from stdin: the order groups
// are dumped in and assignments renumbered to, INCLUDING where its std::sort leaves groups of equal count.
// tests/golden/make_reference_sorted_groupids.py commits what it prints; lbio_sorted_groupids must reproduce it.
//   stdin: n_kinds, then per kind: n_groups counts...
#include <cstdio>
#include <loom/cross_cat.hpp>

int main() {
    int K;
    if (std::scanf("%d", &K) != 1) return 2;
    loom::CrossCat cross_cat;
    for (int k = 0; k < K; ++k) {
        int n;
        if (std::scanf("%d", &n) != 1) return 2;
        loom::CrossCat::Kind &kind = cross_cat.kinds.packed_add();
        auto &counts = kind.mixture.clustering.counts();
        counts.resize(n);
        for (auto &c : counts) if (std::scanf("%d", &c) != 1) return 2;
        kind.mixture.id_tracker.init(n);              // packed == global: a freshly loaded mixture
    }
    const auto sorted = cross_cat.get_sorted_groupids();
    std::printf("[");
    for (int k = 0; k < K; ++k) {
        std::printf("%s[", k ? ", " : "");
        for (size_t i = 0; i < sorted[k].size(); ++i) std::printf("%s%u", i ? "," : "", sorted[k][i]);
        std::printf("]");
    }
    std::printf("]\n");
    return 0;
}
