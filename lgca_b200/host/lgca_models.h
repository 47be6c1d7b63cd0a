// Model descriptors of the host layer: direction count, direction permutations and lattice vectors
// (the data of the reference's ModelDescriptor<M>, src/lgca_models.h:35-52, :228-247, :434-454, :652-672).
// The collision / streaming arithmetic is NOT here: it lives on the device (csrc/lgca_collide.cuh).
#ifndef LGCA_B200_HOST_MODELS_H_
#define LGCA_B200_HOST_MODELS_H_

#include "lgca_common.h"

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

template <Model model_> struct ModelDescriptor;
template <> struct ModelDescriptor<Model::HPP>     : detail::HppTables    { static constexpr int C_ABI_ID = 0; };
template <> struct ModelDescriptor<Model::FHP_I>   : detail::FhpTables<6> { static constexpr int C_ABI_ID = 1; };
template <> struct ModelDescriptor<Model::FHP_II>  : detail::FhpTables<7> { static constexpr int C_ABI_ID = 2; };
template <> struct ModelDescriptor<Model::FHP_III> : detail::FhpTables<7> { static constexpr int C_ABI_ID = 3; };

} // namespace lgca

#endif
