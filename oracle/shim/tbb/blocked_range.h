// TEST INFRASTRUCTURE ONLY (oracle build shim). Minimal tbb::blocked_range stand-in.
#pragma once
#include <cstddef>
namespace tbb {
template <typename T>
class blocked_range {
public:
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end()   const { return e_; }
private:
    T b_, e_;
};
} // namespace tbb
