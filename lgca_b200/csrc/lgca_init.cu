// Device-side set-up for lattices that are too large to initialise through the host:
//   * BC painting straight into the solid mask planes (same geometry as the reference's painters,
//     Lattice<M>::apply_bc_*, src/lattice.cpp:221-334 -- incl. the Karman disc centre (dim_x/6, dim_y/2),
//     diameter float(dim_y/3), float-rounded distance, strict '<', src/lattice.cpp:252-280);
//   * synthetic occupancy with P = 1/NUM_DIR in FLUID cells and chirality with P = 1/2 from a
//     counter-based hash of (seed, global cell, direction) -- the role of Lattice<M>::init_random
//     (src/lattice.cpp:198-217) and Bitset::fill_random (src/lgca_bitset.h:220-224) for >= 1e8-cell
//     throughput runs, where ~8 host rand() calls per cell would take minutes.
//     The hash stream is NOT glibc's; parity runs upload the reference's own initial data instead.
#include "lgca_internal.h"

namespace lgca_b200 {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// one thread per (stored row, word); fills every occupation plane and the chirality plane
template <int ND>
__global__ void __launch_bounds__(128) init_random_kernel(uint32_t* __restrict__ planes, uint32_t* __restrict__ ch,
                                                          const uint32_t* __restrict__ ns,
                                                          const uint32_t* __restrict__ sl, const Geom g, uint64_t seed)
{
    const uint32_t w = blockIdx.y * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.x;
    if (w >= g.nw) return;
    const uint32_t gy = (g.y0 + g.dim_y - g.halo % g.dim_y + y) % g.dim_y;
    const size_t   o  = (size_t)y * g.pitch + w;
    const uint32_t fluid = ~(ns[o] | sl[o]) & valid_mask(g, (int)w);
    const uint32_t thr = (uint32_t)(0x100000000ull / ND);
    uint32_t v[ND], c = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) v[d] = 0;
    for (int b = 0; b < 32; ++b) {
        const uint64_t cell = (uint64_t)gy * g.dim_x + (uint64_t)w * 32u + b;
        const uint64_t hc = mix64(seed ^ (cell * 0xD1B54A32D192ED03ull));
        c |= (uint32_t)(hc >> 63) << b;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const uint32_t u = (uint32_t)(mix64(hc + (uint64_t)(d + 1)) >> 32);
            v[d] |= (u < thr ? 1u : 0u) << b;
        }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) planes[(size_t)d * g.plane_stride + o] = v[d] & fluid;
    ch[o] = c & valid_mask(g, (int)w);
}

// bc_kind: 0 periodic, 1 pipe, 2 karman, 3 reflecting back, 4 reflecting forward
__global__ void __launch_bounds__(128) paint_bc_kernel(uint32_t* __restrict__ ns, uint32_t* __restrict__ sl,
                                                       const Geom g, int bc_kind)
{
    const uint32_t w = blockIdx.y * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.x;
    if (w >= g.nw) return;
    const uint32_t gy = (g.y0 + g.dim_y - g.halo % g.dim_y + y) % g.dim_y;
    const uint32_t vm = valid_mask(g, (int)w);
    uint32_t solid = 0;
    const bool edge_row = (gy == 0 || gy == g.dim_y - 1);
    if (bc_kind >= 1 && edge_row) solid = 0xFFFFFFFFu;
    if (bc_kind >= 3) {
        if (w == 0) solid |= 1u;
        if (w == (g.dim_x - 1) / 32) solid |= 1u << ((g.dim_x - 1) & 31);
    }
    if (bc_kind == 2) {
        const int    cx = (int)(g.dim_x / 6), cy = (int)(g.dim_y / 2);
        const float  diameter = (float)(g.dim_y / 3);
        const double rad = diameter / 2.0;
        for (int b = 0; b < 32; ++b) {
            const double dx = (double)((int)(w * 32u + b) - cx), dy = (double)((int)gy - cy);
            const float  dist = (float)sqrt(dx * dx + dy * dy);
            if ((double)dist < rad) solid |= 1u << b;
        }
    }
    solid &= vm;
    const size_t o = (size_t)y * g.pitch + w;
    ns[o] = (bc_kind == 4) ? 0u : solid;
    sl[o] = (bc_kind == 4) ? solid : 0u;
}

int launch_init_random(lgca_b200_lattice* h, uint32_t* planes, uint64_t seed, cudaStream_t s)
{
    const Geom& g = h->g;
    dim3 grid(g.rows, (g.nw + 127) / 128, 1);
    if (h->nd == 4) init_random_kernel<4><<<grid, 128, 0, s>>>(planes, h->ch, h->ns, h->sl, g, seed);
    else if (h->nd == 6) init_random_kernel<6><<<grid, 128, 0, s>>>(planes, h->ch, h->ns, h->sl, g, seed);
    else init_random_kernel<7><<<grid, 128, 0, s>>>(planes, h->ch, h->ns, h->sl, g, seed);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_paint_bc(lgca_b200_lattice* h, int bc_kind, cudaStream_t s)
{
    const Geom& g = h->g;
    dim3 grid(g.rows, (g.nw + 127) / 128, 1);
    paint_bc_kernel<<<grid, 128, 0, s>>>(h->ns, h->sl, g, bc_kind);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace lgca_b200
