// loom_mix -- the process loom.runner starts for `loom.generate`'s mixing step (src/mix.cc), on the B200 engine:
//
//   loom_mix CONFIG_IN ROWS_IN MODEL_IN GROUPS_IN ASSIGN_IN MODEL_OUT GROUPS_OUT ASSIGN_OUT      (src/mix.cc:32-62)
//
// config.generate.sample_skip sweeps of { REMOVE + ADD every row; kind kernel; hyper kernel } over a fully assigned
// state (Loom::mix, src/loom.cc:564-586), then Loom::dump.
#include <cstdio>
#include <cstdlib>

#include "loom.hpp"

static const char *help_message =
    "Usage: loom_mix CONFIG_IN ROWS_IN MODEL_IN GROUPS_IN ASSIGN_IN\n"
    "  MODEL_OUT GROUPS_OUT ASSIGN_OUT\n"
    "Arguments:\n"
    "  CONFIG_IN     filename of config (e.g. config.pb.gz)\n"
    "  ROWS_IN       filename of input dataset stream (e.g. rows.pbs.gz)\n"
    "  MODEL_IN      filename of input model (e.g. model.pb.gz)\n"
    "  GROUPS_IN     dirname of input per-kind group files\n"
    "  ASSIGN_IN     filename of input assignments stream (e.g. assign.pbs.gz)\n"
    "  MODEL_OUT     filename of output model (e.g. model.pb.gz)\n"
    "  GROUPS_OUT    dirname of output per-kind group files\n"
    "  ASSIGN_OUT    filename of output assignments stream (e.g. assign.pbs.gz)\n"
    "Notes:\n"
    "  Any filename can end with .gz to indicate gzip compression.\n";

int main(int argc, char **argv) {
    if (argc != 9) {
        std::fprintf(stderr, "%s", help_message);
        return 1;
    }
    lbio_config config;
    LOOM_B200_IO(lbio_config_read(argv[1], &config));
    const char *device_env = std::getenv("LOOM_B200_DEVICE");
    loom::b200::Loom engine(config, argv[3], argv[4], argv[5], nullptr, device_env ? std::atoi(device_env) : 0);
    engine.mix(argv[2]);
    engine.dump(argv[6], argv[7], argv[8]);
    return 0;
}
