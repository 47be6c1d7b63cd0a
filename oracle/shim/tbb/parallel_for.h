// TEST INFRASTRUCTURE ONLY (oracle build shim). tbb::parallel_for stand-in:
// the range is cut into chunks that are distributed over OpenMP threads (one chunk = one
// call of the body), so the reference's hot loop uses all host cores when OMP_NUM_THREADS>1
// and runs serially, in order, at OMP_NUM_THREADS=1.
#pragma once
#include "blocked_range.h"
#include <omp.h>
namespace tbb {
template <typename T, typename F>
void parallel_for(const blocked_range<T>& r, const F& f) {
    const T n = r.end() - r.begin();
    const int nthreads = omp_get_max_threads();
    if (nthreads <= 1 || n < (T)1024) { f(r); return; }
    const T nchunks = (T)nthreads * 8;
    const T chunk = (n + nchunks - 1) / nchunks;
#pragma omp parallel for schedule(dynamic, 1)
    for (long long c = 0; c < (long long)nchunks; ++c) {
        T b = r.begin() + (T)c * chunk;
        T e = b + chunk;
        if (b >= r.end()) continue;
        if (e > r.end()) e = r.end();
        f(blocked_range<T>(b, e));
    }
}
} // namespace tbb
