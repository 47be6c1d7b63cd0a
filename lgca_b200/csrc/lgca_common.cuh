// Shared definitions of the device code: bit-plane geometry and periodic word access.
//
// Device layout ("bit-planes", multi-spin coding): for every lattice direction d one plane of
//   rows x pitch 32-bit words; bit j of word w of row y  <=>  site x = 32*w + j of row y.
// Plane d starts at d * plane_stride words.  Bits at x >= dim_x in the last word of a row are kept
// zero ("canonical form").  Static masks (no-slip solids, slip solids, chirality) use the same
// row/word geometry with one plane each.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#include "lgca_collide.cuh"

namespace lgca_b200 {

struct Geom {
    uint32_t dim_x;        // sites per row
    uint32_t dim_y;        // GLOBAL number of rows (periodic torus, src/omp_lattice.cpp:150-176)
    uint32_t nw;           // words per row that hold sites = ceil(dim_x / 32)
    uint32_t rem;          // dim_x % 32 (0 => every word is full)
    uint32_t pitch;        // allocated words per row (>= nw, multiple of 4)
    uint32_t rows;         // rows stored in this handle (strip rows + 2*halo)
    uint32_t halo;         // ghost rows on each side of the strip (0 on a single GPU)
    uint32_t y0;           // global row index of stored row `halo` (first owned row)
    uint32_t wrap_y;       // 1: stored rows are the whole torus (periodic in y inside the kernel)
    uint32_t row_south;    // stored row that is the southern domain edge (global y = 0), or ~0u
    uint32_t row_north;    // stored row that is the northern domain edge (global y = dim_y-1), or ~0u
    uint64_t plane_stride; // words between consecutive planes = rows * pitch
};

__device__ __forceinline__ uint32_t low_mask(int nbits) // nbits in [0,32]
{
    return nbits >= 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u);
}

// mask of the bits of word wi (0 <= wi < nw) that hold real sites
__device__ __forceinline__ uint32_t valid_mask(const Geom& g, int wi)
{
    return (g.rem != 0 && wi == (int)g.nw - 1) ? low_mask((int)g.rem) : 0xFFFFFFFFu;
}

// 32 consecutive sites of a periodic row starting at site 32*wi (wi may be negative or >= nw):
// bit j of the result = site (32*wi + j) mod dim_x.  Full in-range words are a plain load; words that
// touch the row end are assembled bit-exactly from the periodic images (any dim_x >= 1).
__device__ __forceinline__ uint32_t fetch_word(const uint32_t* __restrict__ row, int wi, const Geom& g)
{
    if (g.rem == 0) {
        if (wi < 0) wi += (int)g.nw * ((-wi + (int)g.nw - 1) / (int)g.nw);
        else if (wi >= (int)g.nw) wi %= (int)g.nw;
        return __ldg(row + wi);
    }
    if (wi >= 0 && wi < (int)g.nw - 1) return __ldg(row + wi);
    long long p = ((long long)wi * 32) % (long long)g.dim_x;
    if (p < 0) p += g.dim_x;
    uint32_t r = 0;
    int got = 0;
    while (got < 32) {
        const int w = (int)(p >> 5), o = (int)(p & 31);
        const uint32_t v = __ldg(row + w) >> o;
        int avail = min(32 - o, (int)((long long)g.dim_x - p));
        avail = min(avail, 32 - got);
        r |= (v & low_mask(avail)) << got;
        got += avail;
        p += avail;
        if (p >= (long long)g.dim_x) p = 0;
    }
    return r;
}

// stored-row index of global-relative row y (y may be -1 or rows): periodic when wrap_y, else clamped
// reads of out-of-range rows return row 0 / rows-1 (their results are never used: halo rows are
// re-imported before they matter).
__device__ __forceinline__ uint32_t row_index(int y, const Geom& g)
{
    if (g.wrap_y) {
        if (y < 0) y += (int)g.rows;
        else if (y >= (int)g.rows) y -= (int)g.rows;
        return (uint32_t)y;
    }
    return (uint32_t)max(0, min(y, (int)g.rows - 1));
}

#define LGCA_CUDA_CHECK(expr)                                                                     \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) return lgca_b200::set_cuda_error(_e, #expr, __FILE__, __LINE__);   \
    } while (0)

int set_cuda_error(cudaError_t e, const char* what, const char* file, int line);
int set_error(int code, const char* fmt, ...);

} // namespace lgca_b200
