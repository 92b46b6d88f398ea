// lb_struct.cu -- kind kernel (proposer accumulate, feature x kind scoring,
// block Pitman-Yor, feature moves) and hyper kernel.  [under construction]
#include "lb_internal.h"

void lb_free_proposer(Engine *e) {
    for (auto &p : e->prop) {
        if (p.ints) cudaFree(p.ints);
        if (p.flts) cudaFree(p.flts);
    }
    e->prop.clear();
    e->prop_scores.clear();
}

extern "C" int lb_kind_proposer_score(lb_handle, float *) { return lb_fail("kind_proposer_score: not built yet"); }
extern "C" int lb_kind_run(lb_handle, int32_t, uint64_t, uint32_t *, int32_t *) { return lb_fail("kind_run: not built yet"); }
extern "C" int lb_init_featureless_kinds(lb_handle, int32_t, uint64_t) { return lb_fail("init_featureless_kinds: not built yet"); }
extern "C" int lb_set_hyper_prior(lb_handle, const lb_hyper_prior *) { return lb_fail("set_hyper_prior: not built yet"); }
extern "C" int lb_hyper_run(lb_handle, uint64_t) { return lb_fail("hyper_run: not built yet"); }
