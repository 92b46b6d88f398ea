// loom_posterior_enum -- the process loom's own statistical test drives
// (loom/test/test_posterior_enum.py -> loom.runner.posterior_enum), on the B200 engine: the seven
// positional arguments of src/posterior_enum.cc:57-65 in the same order.
//
//   loom_posterior_enum CONFIG_IN ROWS_IN TARES_IN MODEL_IN GROUPS_IN ASSIGN_IN SAMPLES_OUT
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "loom.hpp"

static const char *help_message =
    "Usage: loom_posterior_enum CONFIG_IN ROWS_IN TARES_IN MODEL_IN GROUPS_IN ASSIGN_IN SAMPLES_OUT\n"
    "Arguments:\n"
    "  CONFIG_IN     config (e.g. config.pb.gz)\n"
    "  ROWS_IN       input dataset stream (e.g. rows.pbs.gz)\n"
    "  TARES_IN      tare rows (e.g. tares.pbs.gz), or --none if the data has not been tared\n"
    "  MODEL_IN      model (e.g. model.pb.gz)\n"
    "  GROUPS_IN     directory of per-kind group files, or --none for empty groups\n"
    "  ASSIGN_IN     assignments stream (e.g. assign.pbs.gz), or --none\n"
    "  SAMPLES_OUT   samples stream to write (e.g. samples.pbs.gz)\n";

int main(int argc, char **argv) {
    if (argc != 8) {
        std::fprintf(stderr, "%s", help_message);
        return 1;
    }
    auto optional_file = [](const char *a) -> const char * { return std::strcmp(a, "--none") == 0 ? nullptr : a; };
    const char *config_in = argv[1];
    const char *rows_in = argv[2];
    const char *tares_in = optional_file(argv[3]);
    const char *model_in = argv[4];
    const char *groups_in = optional_file(argv[5]);
    const char *assign_in = optional_file(argv[6]);
    const char *samples_out = argv[7];

    lbio_config config;
    LOOM_B200_IO(lbio_config_read(config_in, &config));
    const char *device_env = std::getenv("LOOM_B200_DEVICE");
    loom::b200::Loom engine(config, model_in, groups_in, assign_in, tares_in, device_env ? std::atoi(device_env) : 0);
    engine.posterior_enum(rows_in, samples_out);
    return 0;
}
