// TEST INFRASTRUCTURE ONLY (oracle build shim; never shipped in the product path).
// Stand-in for <tbb/spin_mutex.h>: the reference's lgca_bitset.h only names the type in an
// unused typedef, and silently relies on <limits>/<cstdint> arriving through this include.
#pragma once
#include <limits>
#include <cstdint>
#include <cstddef>
namespace tbb { struct spin_mutex {}; }
