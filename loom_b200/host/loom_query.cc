// loom_query -- the process loom.runner starts for `loom.query` (src/query.cc), on the B200 engine:
// same five positional arguments (src/query.cc:34-62, loom/runner.py:362-368), same request /
// response streams, abort() with loom's error text on failure.
//
//   loom_query ROOT_IN REQUESTS_IN CONFIG_IN RESPONSES_OUT LOG_OUT
//
// `-` / `-.gz` = stdin / stdout (the piped server of loom/runner.py:376-381); `--none` = no log.
// LOOM_B200_DEVICE selects the GPU (default 0).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "query_server.hpp"

static const char *help_message =
    "Usage: loom_query ROOT_IN REQUESTS_IN CONFIG_IN RESPONSES_OUT LOG_OUT\n"
    "Arguments:\n"
    "  ROOT_IN         root dirname of dataset in loom store\n"
    "  REQUESTS_IN     filename of requests stream (e.g. requests.pbs.gz)\n"
    "  CONFIG_IN       filename of query config (e.g. config.pb.gz)\n"
    "  RESPONSES_OUT   filename of responses stream (e.g. responses.pbs.gz)\n"
    "  LOG_OUT         filename of log (e.g. log.pbs.gz)\n"
    "                  or --none to not log\n"
    "Notes:\n"
    "  Any filename can end with .gz to indicate gzip compression.\n"
    "  Any filename can be '-' or '-.gz' to indicate stdin/stdout.\n";

int main(int argc, char **argv) {
    if (argc != 6) {
        std::fprintf(stderr, "%s", help_message);
        return 1;
    }
    const char *root_in = argv[1];
    const char *requests_in = argv[2];
    const char *config_in = argv[3];
    const char *responses_out = argv[4];
    const char *log_out = std::strcmp(argv[5], "--none") == 0 ? nullptr : argv[5];

    lbio_log *log = nullptr;
    if (log_out) LOOM_B200_IO(lbio_log_open(log_out, 1, &log));   // loom::logger.append (src/query.cc:62-64)

    lbio_config config;
    LOOM_B200_IO(lbio_config_read(config_in, &config));
    const char *device_env = std::getenv("LOOM_B200_DEVICE");
    loom::b200::QueryServer server(root_in, config, device_env ? std::atoi(device_env) : 0);
    server.serve(requests_in, responses_out);
    if (log) lbio_log_close(log);
    return 0;
}
