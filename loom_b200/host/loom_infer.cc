// loom_infer -- the process loom.runner starts for `loom.tasks.infer` (src/infer.cc), on the
// B200 engine: same twelve positional arguments in the same order (src/infer.cc:68-81,
// loom/runner.py:238-243), same files in and out, abort() with loom's error text on failure.
//
//   loom_infer CONFIG_IN ROWS_IN TARES_IN MODEL_IN GROUPS_IN ASSIGN_IN CHECKPOINT_IN
//              MODEL_OUT GROUPS_OUT ASSIGN_OUT CHECKPOINT_OUT LOG_OUT
//
// `--none` = absent; names ending in .gz are gzip-compressed (src/args.hpp:56-64).
// LOOM_B200_DEVICE selects the GPU (default 0).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "loom.hpp"

static const char *help_message =
    "Usage: loom_infer CONFIG_IN ROWS_IN TARES_IN MODEL_IN GROUPS_IN ASSIGN_IN CHECKPOINT_IN\n"
    "  MODEL_OUT GROUPS_OUT ASSIGN_OUT CHECKPOINT_OUT LOG_OUT\n"
    "Arguments:\n"
    "  CONFIG_IN         config (e.g. config.pb.gz)\n"
    "  ROWS_IN           input dataset stream (e.g. rows.pbs.gz)\n"
    "  TARES_IN          tare rows (e.g. tares.pbs.gz), or --none if the data has not been tared\n"
    "  MODEL_IN          model (e.g. model.pb.gz)\n"
    "  GROUPS_IN         directory of per-kind group files, or --none for empty groups\n"
    "  ASSIGN_IN         assignments stream (e.g. assign.pbs.gz), or --none\n"
    "  CHECKPOINT_IN     checkpoint state (e.g. checkpoint.pb.gz), or --none\n"
    "  MODEL_OUT         model to write, or --none\n"
    "  GROUPS_OUT        directory for per-kind group files, or --none\n"
    "  ASSIGN_OUT        assignments stream to write, or --none\n"
    "  CHECKPOINT_OUT    checkpoint state to write, or --none\n"
    "  LOG_OUT           log stream (e.g. log.pbs.gz), or --none\n"
    "Notes:\n"
    "  Any filename can end with .gz to indicate gzip compression.\n";

int main(int argc, char **argv) {
    if (argc != 13) {
        std::fprintf(stderr, "%s", help_message);
        return 1;
    }
    int next = 1;
    auto pop = [&]() -> const char * { return argv[next++]; };
    auto pop_optional_file = [&]() -> const char * {   // Args::pop_optional_file (src/args.hpp:56-64)
        const char *a = argv[next++];
        return std::strcmp(a, "--none") == 0 ? nullptr : a;
    };
    const char *config_in = pop();
    const char *rows_in = pop();
    const char *tares_in = pop_optional_file();
    const char *model_in = pop();
    const char *groups_in = pop_optional_file();
    const char *assign_in = pop_optional_file();
    const char *checkpoint_in = pop_optional_file();
    const char *model_out = pop_optional_file();
    const char *groups_out = pop_optional_file();
    const char *assign_out = pop_optional_file();
    const char *checkpoint_out = pop_optional_file();
    const char *log_out = pop_optional_file();

    // LOOM_B200_TIMING=1: wall clock of the process's phases on stderr (where a run's seconds go: DESIGN.md 9)
    const bool timing = std::getenv("LOOM_B200_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[LOOM_B200_TIMING] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    lbio_config config;
    LOOM_B200_IO(lbio_config_read(config_in, &config));
    const char *device_env = std::getenv("LOOM_B200_DEVICE");
    loom::b200::Loom engine(config, model_in, groups_in, assign_in, tares_in, device_env ? std::atoi(device_env) : 0);
    if (log_out) engine.log_to(log_out, checkpoint_in != nullptr);   // append when resuming (src/infer.cc:83-89)
    lap("context + model + groups");

    if (config.extra_passes > 0) {
        engine.infer_multi_pass(rows_in, checkpoint_in, checkpoint_out);
        lap("infer_multi_pass");
        engine.dump(model_out, groups_out, assign_out);
        lap("dump");
    } else {
        engine.infer_single_pass(rows_in, assign_out);
        lap("infer_single_pass");
        engine.dump(model_out, groups_out);
        lap("dump");
    }
    return 0;
}
