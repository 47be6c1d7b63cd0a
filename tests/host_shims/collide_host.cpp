// Test shim: exposes the bit-sliced collision / wall networks of lgca_b200/csrc/lgca_collide.cuh
// (the very code the CUDA kernels inline) to the CPU unit tests through their host code path.
#include "../../lgca_b200/csrc/lgca_collide.cuh"

using namespace lgca_b200;

template <int MODEL>
static void run(uint32_t* n, uint32_t p, uint32_t ns, uint32_t sl, uint32_t ew, uint32_t ns_row)
{
    uint32_t v[7];
    for (int d = 0; d < 7; ++d) v[d] = n[d];
    collide_and_walls<MODEL, true, true>(v, p, ns, sl, ew, ns_row);
    for (int d = 0; d < 7; ++d) n[d] = v[d];
}

extern "C" void lgca_host_collide_words(int model, uint32_t* n, uint32_t p, uint32_t ns, uint32_t sl, uint32_t ew,
                                        uint32_t ns_row)
{
    switch (model) {
    case 0: run<MODEL_HPP>(n, p, ns, sl, ew, ns_row); break;
    case 1: run<MODEL_FHP_I>(n, p, ns, sl, ew, ns_row); break;
    case 2: run<MODEL_FHP_II>(n, p, ns, sl, ew, ns_row); break;
    default: run<MODEL_FHP_III>(n, p, ns, sl, ew, ns_row); break;
    }
}
