// loom_sparsify -- the process loom.runner starts for the sparsify pass (src/sparsify.cc), on the B200 engine:
//
//   loom_sparsify SCHEMA_ROW_IN TARES_IN ROWS_IN ROWS_OUT           (src/sparsify.cc:32-53)
//
// Every row relative to the tare row: ProductValue::Diff{pos, neg, tares: [0]} (src/differ.cc:148-189, 248-298;
// worked example doc/adapting.md:323-414).  With no tare the rows pass through.
#include <cstdio>
#include <cstdlib>

#include "differ.hpp"

static const char *help_message =
    "Usage: loom_sparsify SCHEMA_ROW_IN TARES_IN ROWS_IN ROWS_OUT\n"
    "Arguments:\n"
    "  SCHEMA_ROW_IN filename of schema row (e.g. schema.pb.gz)\n"
    "  TARES_IN      filename of tare rows (e.g. tares.pbs.gz)\n"
    "  ROWS_IN       filename of input dataset stream (e.g. rows.pbs.gz)\n"
    "  ROWS_OUT      filename of output dataset stream (e.g. diffs.pbs.gz)\n"
    "Notes:\n"
    "  Any filename can end with .gz to indicate gzip compression.\n";

int main(int argc, char **argv) {
    if (argc != 5) {
        std::fprintf(stderr, "%s", help_message);
        return 1;
    }
    const char *device_env = std::getenv("LOOM_B200_DEVICE");
    loom::b200::Differ differ(argv[1], device_env ? std::atoi(device_env) : 0);
    differ.tares_load(argv[2]);
    differ.compress_rows(argv[3], argv[4]);
    return 0;
}
