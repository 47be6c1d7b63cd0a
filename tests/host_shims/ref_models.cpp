// Test shim (only where /root/reference exists): the REFERENCE's ModelDescriptor<M> offset tables, for comparison.
#include <cmath>
#include "lgca_common.h"
#include "lgca_models.h"

using namespace lgca;

template <Model M>
static void offsets(unsigned dx, unsigned dy, int* out)
{
    ModelDescriptor<M> m(dx, dy);
    const int* t[10] = {m.offset_to_neighbor_even, m.offset_to_neighbor_odd, m.offset_to_eastern_boundary_even,
                        m.offset_to_eastern_boundary_odd, m.offset_to_northern_boundary_even, m.offset_to_northern_boundary_odd,
                        m.offset_to_western_boundary_even, m.offset_to_western_boundary_odd, m.offset_to_southern_boundary_even,
                        m.offset_to_southern_boundary_odd};
    for (int a = 0; a < 10; ++a)
        for (unsigned d = 0; d < 7; ++d) out[a * 7 + d] = d < ModelDescriptor<M>::NUM_DIR ? t[a][d] : 0;
}

extern "C" void lgca_ref_model_offsets(int model, unsigned dx, unsigned dy, int* out)
{
    switch (model) {
    case 0: offsets<Model::HPP>(dx, dy, out); break;
    case 1: offsets<Model::FHP_I>(dx, dy, out); break;
    case 2: offsets<Model::FHP_II>(dx, dy, out); break;
    default: offsets<Model::FHP_III>(dx, dy, out); break;
    }
}
