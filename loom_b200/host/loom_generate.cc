// loom_generate -- the process loom.runner starts for `loom.generate` (src/generate.cc), on the B200 engine:
//
//   loom_generate CONFIG_IN MODEL_IN ROWS_OUT MODEL_OUT GROUPS_OUT ASSIGN_OUT            (src/generate.cc:32-62)
//
// config.generate.row_count rows synthesised from the model (Loom::generate, src/loom.cc:548-562): hypers from the
// prior grids, a Pitman-Yor seating per kind, values from the groups' posterior predictives at config.generate.density.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "loom.hpp"

static const char *help_message =
    "Usage: loom_generate CONFIG_IN MODEL_IN ROWS_OUT MODEL_OUT GROUPS_OUT ASSIGN_OUT\n"
    "Arguments:\n"
    "  CONFIG_IN     filename of config (e.g. config.pb.gz)\n"
    "  MODEL_IN      filename of model (e.g. model.pb.gz)\n"
    "  ROWS_OUT      filename of output dataset stream (e.g. rows.pbs.gz)\n"
    "  MODEL_OUT     filename of model to write, or --none to discard model\n"
    "  GROUPS_OUT    dirname to contain per-kind group files\n"
    "                or --none to discard groups\n"
    "  ASSIGN_OUT    filename of assignments stream (e.g. assign.pbs.gz)\n"
    "                or --none to discard assignments\n"
    "Notes:\n"
    "  Any filename can end with .gz to indicate gzip compression.\n";

int main(int argc, char **argv) {
    if (argc != 7) {
        std::fprintf(stderr, "%s", help_message);
        return 1;
    }
    auto optional = [](const char *a) -> const char * { return std::strcmp(a, "--none") == 0 ? nullptr : a; };
    lbio_config config;
    LOOM_B200_IO(lbio_config_read(argv[1], &config));
    const char *device_env = std::getenv("LOOM_B200_DEVICE");
    loom::b200::Loom engine(config, argv[2], nullptr, nullptr, nullptr, device_env ? std::atoi(device_env) : 0);
    engine.generate(argv[3]);
    engine.dump(optional(argv[4]), optional(argv[5]), optional(argv[6]));
    return 0;
}
