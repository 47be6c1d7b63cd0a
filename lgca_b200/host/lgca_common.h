// Basic types of the host layer.  Same names and values as the reference's src/lgca_common.h:46-60
// (they are part of the Lattice<Model> API every app is written against).
#ifndef LGCA_B200_HOST_COMMON_H_
#define LGCA_B200_HOST_COMMON_H_

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

namespace lgca {

using std::string;

typedef float Real; // floating-point precision of all physics scalars and fields

enum class Model { HPP, FHP_I, FHP_II, FHP_III };

enum class CellType : int { FLUID = 0, SOLID_NO_SLIP = 1, SOLID_SLIP = 2 };

} // namespace lgca

#endif
