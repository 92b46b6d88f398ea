// TEST INFRASTRUCTURE.  Drives the REFERENCE'S OWN loom::Differ (src/differ.cc + src/product_value.cc, compiled
// unmodified from /root/reference by `make -C oracle ref`; oracle/ref_stubs_value/ stands in for protoc's generated
// classes, libprotobuf's streams and the un-vendored distributions headers) over a table of rows read from stdin:
// Differ::add_rows (the tare pass, what `loom tare` runs) and Differ::compress_rows (the sparsify pass), and prints the
// tare row and every compressed row as one JSON document.  tests/golden/make_reference_differ.py turns that into the
// committed fixture the product's differ (lb_differ_*, lbio_rows_writer_write) and the oracle's must reproduce.
//   stdin: booleans_size counts_size reals_size n_rows, then n_rows lines of one token per feature ('-' = unobserved)
#include <cstdio>
#include <cstring>
#include <string>
#include <loom/differ.hpp>

static void print_value(const loom::ValueSchema &schema, const loom::ProductValue &value) {
    typedef loom::ProductValue::Observed Observed;
    const Observed &o = value.observed();
    std::printf("{\"sparsity\": \"%s\", \"features\": [", Observed::Sparsity_Name(o.sparsity()));
    bool first = true;
    const size_t size = schema.total_size();
    switch (o.sparsity()) {
        case Observed::ALL: for (size_t i = 0; i < size; ++i) { std::printf("%s%zu", first ? "" : ",", i); first = false; } break;
        case Observed::DENSE: for (size_t i = 0; i < size; ++i) if (o.dense(i)) { std::printf("%s%zu", first ? "" : ",", i); first = false; } break;
        case Observed::SPARSE: for (int i = 0; i < o.sparse_size(); ++i) { std::printf("%s%u", first ? "" : ",", o.sparse(i)); first = false; } break;
        case Observed::NONE: break;
    }
    std::printf("], \"booleans\": [");
    for (int i = 0; i < value.booleans_size(); ++i) std::printf("%s%d", i ? "," : "", (int)value.booleans(i));
    std::printf("], \"counts\": [");
    for (int i = 0; i < value.counts_size(); ++i) std::printf("%s%u", i ? "," : "", value.counts(i));
    std::printf("], \"reals\": [");
    for (int i = 0; i < value.reals_size(); ++i) std::printf("%s%.9g", i ? "," : "", value.reals(i));
    std::printf("]}");
}

int main() {
    loom::ValueSchema schema;
    size_t n_rows;
    if (std::scanf("%zu %zu %zu %zu", &schema.booleans_size, &schema.counts_size, &schema.reals_size, &n_rows) != 4) return 2;
    const size_t size = schema.total_size();
    auto &rows = loom::protobuf::memory_files()["rows"];
    char token[64];
    for (size_t r = 0; r < n_rows; ++r) {
        loom::protobuf::Row row;
        row.set_id(r);
        loom::ProductValue &value = *row.mutable_diff()->mutable_pos();
        value.mutable_observed()->set_sparsity(loom::ProductValue::Observed::DENSE);
        row.mutable_diff()->mutable_neg()->mutable_observed()->set_sparsity(loom::ProductValue::Observed::NONE);
        for (size_t i = 0; i < size; ++i) {
            if (std::scanf("%63s", token) != 1) return 2;
            const bool observed = std::strcmp(token, "-") != 0;
            value.mutable_observed()->add_dense(observed);
            if (!observed) continue;
            if (i < schema.booleans_size) value.add_booleans(std::atoi(token) != 0);
            else if (i < schema.booleans_size + schema.counts_size) value.add_counts((uint32_t)std::strtoul(token, nullptr, 10));
            else value.add_reals(std::strtof(token, nullptr));
        }
        rows.push_back(row);
    }

    loom::Differ differ(schema);
    differ.add_rows("rows");                       // `loom_tare`: src/tare.cc
    const loom::ProductValue tare = differ.get_tare();
    differ.compress_rows("rows", "diffs");         // `loom_sparsify`: src/sparsify.cc (a Differ with this tare)

    std::printf("{\"tare\": ");
    print_value(schema, tare);
    std::printf(", \"rows\": [\n");
    const auto &diffs = loom::protobuf::memory_files()["diffs"];
    for (size_t r = 0; r < diffs.size(); ++r) {
        std::printf("%s{\"id\": %llu, \"pos\": ", r ? ",\n" : "", (unsigned long long)diffs[r].id());
        print_value(schema, diffs[r].diff().pos());
        std::printf(", \"neg\": ");
        print_value(schema, diffs[r].diff().neg());
        std::printf(", \"tares\": %d}", diffs[r].diff().tares_size());
    }
    std::printf("\n]}\n");
    return 0;
}
