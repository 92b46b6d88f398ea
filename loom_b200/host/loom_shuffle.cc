// loom_shuffle -- the process loom.runner starts to shuffle a dataset before inference (src/shuffle.cc):
//
//   loom_shuffle ROWS_IN ROWS_OUT [SEED=0] [TARGET_MEM_BYTES=4e9]                       (src/shuffle.cc:31-58)
//
// Host I/O only (libloomb200_io.so): a seeded permutation of the stream's messages, chunked to the memory target.
#include <cstdio>
#include <cstdlib>

#include "../../include/loom_b200_io.h"

static const char *help_message =
    "Usage: loom_shuffle ROWS_IN ROWS_OUT [SEED=0] [TARGET_MEM_BYTES=4e9]\n"
    "Arguments:\n"
    "  ROWS_IN           filename of input dataset stream (e.g. rows.pbs.gz)\n"
    "  ROWS_OUT          filename of output dataset stream (e.g. rows_out.pbs.gz)\n"
    "  SEED              random seed\n"
    "  TARGET_MEM_BYTES  target memory usage in bytes\n"
    "Notes:\n"
    "  Any filename can end with .gz to indicate gzip compression.\n";

int main(int argc, char **argv) {
    if (argc < 3 || argc > 5) {
        std::fprintf(stderr, "%s", help_message);
        return 1;
    }
    const long seed = argc > 3 ? std::atol(argv[3]) : 0L;
    const double target_mem_bytes = argc > 4 ? std::atof(argv[4]) : 4e9;
    if (lbio_shuffle_stream(argv[1], argv[2], (uint64_t)seed, target_mem_bytes) != 0) {
        std::fprintf(stderr, "ERROR %s\n\t%s : %d\n\t%s\n", lbio_last_error(), __FILE__, __LINE__, __func__);
        std::abort();
    }
    return 0;
}
