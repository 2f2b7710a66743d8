// CPU harness around agarai_b200/csrc/agar_host_expand.h (test infrastructure): the same functions agar_step_host's workers
// run, callable from pytest through ctypes.  Built by tests/test_host_expand.py with g++.
#include "../agarai_b200/csrc/agar_host_expand.h"

using namespace agar;

extern "C" {

// dst <- the dense [C,G,G] image described by (entries, count, meta = {cx, cy, has cells, -})
void he_expand(int G, int C, int K, int obs_flags, float arena, const float* oob_off, float* dst, const ObsEntry* ent, int count,
               const float* meta) {
    const GridGeo geo{G, C, K, obs_flags, arena, oob_off};
    expand_agent(geo, dst, ent, count, meta);
}

// dst holds the previous image (pe, pc, pmeta); afterwards the one described by (ne, nc, nmeta)
void he_delta(int G, int C, int K, int obs_flags, float arena, const float* oob_off, float* dst, const ObsEntry* pe, int pc,
              const float* pmeta, const ObsEntry* ne, int nc, const float* nmeta) {
    const GridGeo geo{G, C, K, obs_flags, arena, oob_off};
    DeltaPlan d;
    delta_plan(geo, pmeta, nmeta, sentinel_mask(geo), d);
    delta_prefetch(geo, dst, d, pe, pc, ne, nc);
    delta_apply(geo, dst, d, pe, pc, ne, nc, nmeta);
}

}  // extern "C"
