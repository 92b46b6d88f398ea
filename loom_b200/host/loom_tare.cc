// loom_tare -- the process loom.runner starts for the tare pass (src/tare.cc), on the B200 engine:
//
//   loom_tare SCHEMA_ROW_IN ROWS_IN TARES_OUT                       (src/tare.cc:31-50)
//
// Per boolean / count column the mode of the observed values; a column whose mode covers more than half of
// all rows goes into the tare row (src/differ.cc:93-146, 191-206).  TARES_OUT holds that one row, or nothing.
#include <cstdio>
#include <cstdlib>

#include "differ.hpp"

static const char *help_message =
    "Usage: loom_tare SCHEMA_ROW_IN ROWS_IN TARES_OUT\n"
    "Arguments:\n"
    "  SCHEMA_ROW_IN filename of schema row (e.g. schema.pb.gz)\n"
    "  ROWS_IN       filename of input dataset stream (e.g. rows.pbs.gz)\n"
    "  TARES_OUT     filename of output tare rows (e.g. tares.pbs.gz)\n"
    "Notes:\n"
    "  Any filename can end with .gz to indicate gzip compression.\n";

int main(int argc, char **argv) {
    if (argc != 4) {
        std::fprintf(stderr, "%s", help_message);
        return 1;
    }
    const char *device_env = std::getenv("LOOM_B200_DEVICE");
    loom::b200::Differ differ(argv[1], device_env ? std::atoi(device_env) : 0);
    differ.add_rows(argv[2]);
    differ.tares_dump(argv[3]);
    return 0;
}
