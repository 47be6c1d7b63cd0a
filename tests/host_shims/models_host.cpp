// Test shim: the host ModelDescriptor<M> of lgca_b200/host/lgca_models.h behind a C interface.
#include "../../lgca_b200/host/lgca_models.h"

using namespace lgca;

template <Model M>
static void offsets(unsigned dx, unsigned dy, int* out) // [10][7]
{
    ModelDescriptor<M> m(dx, dy);
    const int* t[10] = {m.offset_to_neighbor_even, m.offset_to_neighbor_odd, m.offset_to_eastern_boundary_even,
                        m.offset_to_eastern_boundary_odd, m.offset_to_northern_boundary_even, m.offset_to_northern_boundary_odd,
                        m.offset_to_western_boundary_even, m.offset_to_western_boundary_odd, m.offset_to_southern_boundary_even,
                        m.offset_to_southern_boundary_odd};
    for (int a = 0; a < 10; ++a)
        for (unsigned d = 0; d < 7; ++d) out[a * 7 + d] = d < ModelDescriptor<M>::NUM_DIR ? t[a][d] : 0;
}

template <Model M>
static void rule(int what, unsigned char* in, unsigned char* out, int p)
{
    if (what == 0) ModelDescriptor<M>::collide(in, out, p != 0);
    else if (what == 1) ModelDescriptor<M>::bounce_back(in, out);
    else if (what == 2) ModelDescriptor<M>::bounce_forward_x(in, out);
    else ModelDescriptor<M>::bounce_forward_y(in, out);
}

extern "C" void lgca_host_model_offsets(int model, unsigned dx, unsigned dy, int* out)
{
    switch (model) {
    case 0: offsets<Model::HPP>(dx, dy, out); break;
    case 1: offsets<Model::FHP_I>(dx, dy, out); break;
    case 2: offsets<Model::FHP_II>(dx, dy, out); break;
    default: offsets<Model::FHP_III>(dx, dy, out); break;
    }
}

extern "C" void lgca_host_model_rule(int model, int what, unsigned char* in, unsigned char* out, int p)
{
    switch (model) {
    case 0: rule<Model::HPP>(what, in, out, p); break;
    case 1: rule<Model::FHP_I>(what, in, out, p); break;
    case 2: rule<Model::FHP_II>(what, in, out, p); break;
    default: rule<Model::FHP_III>(what, in, out, p); break;
    }
}
