// Model descriptors of the host layer -- the API of the reference's ModelDescriptor<M> (src/lgca_models.h):
//   * tables: NUM_DIR, INV_DIR, MIR_DIR_X/Y, LATTICE_VEC_X/Y (:35-52, :228-247, :434-454, :652-672);
//   * per-cell rules on NUM_DIR 0/1 bytes: collide(in, out, p), bounce_back, bounce_forward_x/y (:134-182, :367-427,
//     :568-613, :786-856).  collide() runs the SAME boolean network the CUDA kernels inline (csrc/lgca_collide.cuh,
//     compiled here for the host with one site per word), so host tools and the device cannot drift apart;
//   * the ten streaming offset tables + constructor(dim_x, dim_y) (:79-132, :292-365, :483-566, :701-784), GENERATED
//     here from the lattice geometry: offset to the neighbour a particle moving in direction d reaches, and the
//     corrections that wrap it around the periodic torus.  One dead entry of the reference is not reproduced: its
//     even-row `offset_to_southern_boundary_even[2]` carries a stray +1 (src/lgca_models.h:328), which no lattice can
//     reach (the top row of an even-height lattice is odd).
// The reference's COLLISION_LUT / BB_LUT / BF_*_LUT byte tables are unused there (most are empty "TODO" arrays) and
// are not provided.  The hot path does not go through this file: the device streams with funnel shifts and row
// arithmetic (csrc/lgca_step_wave.cu), which tests/test_host_models.py checks against these tables.
#ifndef LGCA_B200_HOST_MODELS_H_
#define LGCA_B200_HOST_MODELS_H_

#include "lgca_common.h"
#include "../csrc/lgca_collide.cuh"

namespace lgca {

namespace detail {
// float(sin(pi/3)), the reference's `static constexpr Real SIN`
constexpr Real kSin60 = 0.866025388f;

struct HppTables {
    static constexpr unsigned int NUM_DIR = 4;
    static constexpr char INV_DIR[4]   = {2, 3, 0, 1};
    static constexpr char MIR_DIR_X[4] = {0, 3, 2, 1};
    static constexpr char MIR_DIR_Y[4] = {2, 1, 0, 3};
    static constexpr Real LATTICE_VEC_X[4] = {1.0f, 0.0f, -1.0f, 0.0f};
    static constexpr Real LATTICE_VEC_Y[4] = {0.0f, 1.0f, 0.0f, -1.0f};
};

template <unsigned int N> // N = 6 (FHP-I) or 7 (FHP-II/III: direction 6 is the rest particle)
struct FhpTables {
    static constexpr unsigned int NUM_DIR = N;
    static constexpr Real SIN = kSin60;
    static constexpr char INV_DIR[7]   = {3, 4, 5, 0, 1, 2, 6};
    static constexpr char MIR_DIR_X[7] = {0, 5, 4, 3, 2, 1, 6};
    static constexpr char MIR_DIR_Y[7] = {3, 2, 1, 0, 5, 4, 6};
    static constexpr Real LATTICE_VEC_X[7] = {1.0f, 0.5f, -0.5f, -1.0f, -0.5f, 0.5f, 0.0f};
    static constexpr Real LATTICE_VEC_Y[7] = {0.0f, kSin60, kSin60, 0.0f, -kSin60, -kSin60, 0.0f};
};
} // namespace detail

// Rules + offset tables on top of a table set T; ID = model number of the C-ABI (== lgca_b200::MODEL_*).
template <typename T, int ID>
struct ModelRules : T {
    static constexpr int C_ABI_ID = ID;
    using T::NUM_DIR;

    // Memory offsets (in cells) for the propagation step, per row parity, and their periodic corrections; same
    // names and meaning as the reference's members.  `offset_to_<X>_boundary` is the correction that lands on the
    // boundary X, i.e. it is added when the cell sits on the OPPOSITE edge (src/omp_lattice.cpp:150-176).
    int offset_to_neighbor_even[NUM_DIR], offset_to_neighbor_odd[NUM_DIR];
    int offset_to_eastern_boundary_even[NUM_DIR], offset_to_eastern_boundary_odd[NUM_DIR];
    int offset_to_northern_boundary_even[NUM_DIR], offset_to_northern_boundary_odd[NUM_DIR];
    int offset_to_western_boundary_even[NUM_DIR], offset_to_western_boundary_odd[NUM_DIR];
    int offset_to_southern_boundary_even[NUM_DIR], offset_to_southern_boundary_odd[NUM_DIR];

    // cell displacement of a particle moving in direction d from a row of the given parity (hexagonal rows: even rows
    // sit half a cell to the left of odd rows; HPP is a square lattice)
    static void displacement(unsigned int d, bool odd_row, int& dx, int& dy)
    {
        if (NUM_DIR == 4) { dx = d == 0 ? 1 : (d == 2 ? -1 : 0); dy = d == 1 ? 1 : (d == 3 ? -1 : 0); return; }
        if (d == 6) { dx = dy = 0; return; }
        dy = (d == 1 || d == 2) ? 1 : ((d == 4 || d == 5) ? -1 : 0);
        if (d == 0) dx = 1;
        else if (d == 3) dx = -1;
        else if (d == 1 || d == 5) dx = odd_row ? 1 : 0;    // NE / SE
        else dx = odd_row ? 0 : -1;                          // NW / SW
    }

    ModelRules(const unsigned int dim_x, const unsigned int dim_y)
    {
        const int nx = (int)dim_x, cells = (int)(dim_x * dim_y); // int, like the reference (overflows at >= 2^31 cells)
        for (unsigned int d = 0; d < NUM_DIR; ++d) {
            for (int odd = 0; odd < 2; ++odd) {
                int dx, dy;
                displacement(d, odd != 0, dx, dy);
                (odd ? offset_to_neighbor_odd : offset_to_neighbor_even)[d]                   = dy * nx + dx;
                (odd ? offset_to_eastern_boundary_odd : offset_to_eastern_boundary_even)[d]   = dx < 0 ? nx : 0;      // leaving over the western edge
                (odd ? offset_to_western_boundary_odd : offset_to_western_boundary_even)[d]   = dx > 0 ? -nx : 0;     // ... the eastern edge
                (odd ? offset_to_northern_boundary_odd : offset_to_northern_boundary_even)[d] = dy < 0 ? cells : 0;   // ... the southern edge
                (odd ? offset_to_southern_boundary_odd : offset_to_southern_boundary_even)[d] = dy > 0 ? -cells : 0;  // ... the northern edge
            }
        }
    }

    // collision of one cell: NUM_DIR bytes of 0/1 in, same out; p = the cell's chirality bit
    static inline void collide(unsigned char* node_state_in, unsigned char* node_state_out, const bool p)
    {
        uint32_t n[7] = {0, 0, 0, 0, 0, 0, 0};
        for (unsigned int d = 0; d < NUM_DIR; ++d) n[d] = node_state_in[d] ? 1u : 0u;
        lgca_b200::collide<ID>(n, p ? 1u : 0u);   // the device's network, one site in bit 0
        for (unsigned int d = 0; d < NUM_DIR; ++d) node_state_out[d] = (unsigned char)(n[d] & 1u);
    }
    static inline void bounce_back(unsigned char* node_state_in, unsigned char* node_state_out)
    {
        for (unsigned int d = 0; d < NUM_DIR; ++d) node_state_out[d] = node_state_in[(int)T::INV_DIR[d]];
    }
    static inline void bounce_forward_x(unsigned char* node_state_in, unsigned char* node_state_out)
    {
        for (unsigned int d = 0; d < NUM_DIR; ++d) node_state_out[d] = node_state_in[(int)T::MIR_DIR_X[d]];
    }
    static inline void bounce_forward_y(unsigned char* node_state_in, unsigned char* node_state_out)
    {
        for (unsigned int d = 0; d < NUM_DIR; ++d) node_state_out[d] = node_state_in[(int)T::MIR_DIR_Y[d]];
    }
};

template <Model model_> struct ModelDescriptor;
template <> struct ModelDescriptor<Model::HPP>     : ModelRules<detail::HppTables, 0>    { using ModelRules::ModelRules; };
template <> struct ModelDescriptor<Model::FHP_I>   : ModelRules<detail::FhpTables<6>, 1> { using ModelRules::ModelRules; };
template <> struct ModelDescriptor<Model::FHP_II>  : ModelRules<detail::FhpTables<7>, 2> { using ModelRules::ModelRules; };
template <> struct ModelDescriptor<Model::FHP_III> : ModelRules<detail::FhpTables<7>, 3> { using ModelRules::ModelRules; };

} // namespace lgca

#endif
