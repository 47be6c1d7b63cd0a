// Conversions between the reference's host layouts and the device bit-planes.
//
//   state bytes  (bit d of byte `cell` = direction d; src/omp_lattice.cpp:179,237)   <-> NUM_DIR planes
//   CellType int32 per cell (src/lgca_common.h:53-57)                               ->  no-slip / slip masks
//   chirality Bitset, one flat LSB-first bit-field over all cells
//   (src/lgca_bitset.h:220-231, read as m_rnd_cpu[cell] at src/omp_lattice.cpp:198) ->  chirality plane
//
// One warp transposes 32 sites at a time with __ballot_sync (bytes -> plane words) or shuffles
// (plane words -> bytes).  Row starts need not be aligned to anything: cell = row*dim_x + x.
#include "lgca_internal.h"

namespace lgca_b200 {

// grid: (row, word-group); block = 128 threads = 4 warps, each warp packs one word per iteration
template <int ND>
__global__ void __launch_bounds__(128) pack_state_kernel(const uint8_t* __restrict__ bytes, uint32_t* __restrict__ planes,
                                                         const Geom g, uint32_t row0)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.y * blockDim.x + threadIdx.x) >> 5;
    const uint32_t r = blockIdx.x;              // row inside the chunk
    const uint32_t y = row0 + r;                // stored row
    const int wpb = (gridDim.y * blockDim.x) >> 5;
    for (int w = warp; w < (int)g.nw; w += wpb) {
        const uint32_t x = (uint32_t)w * 32u + lane;
        uint32_t b = 0;
        if (x < g.dim_x) b = bytes[(size_t)r * g.dim_x + x];
        uint32_t mine = 0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, (b >> d) & 1u);
            if (lane == d) mine = word;
        }
        if (lane < ND) planes[(size_t)lane * g.plane_stride + (size_t)y * g.pitch + w] = mine;
    }
}

template <int ND>
__global__ void __launch_bounds__(128) unpack_state_kernel(const uint32_t* __restrict__ planes, uint8_t* __restrict__ bytes,
                                                           const Geom g, uint32_t row0)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.y * blockDim.x + threadIdx.x) >> 5;
    const uint32_t r = blockIdx.x;
    const uint32_t y = row0 + r;
    const int wpb = (gridDim.y * blockDim.x) >> 5;
    for (int w = warp; w < (int)g.nw; w += wpb) {
        uint32_t mine = 0;
        if (lane < ND) mine = planes[(size_t)lane * g.plane_stride + (size_t)y * g.pitch + w];
        uint32_t b = 0;
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const uint32_t word = __shfl_sync(0xFFFFFFFFu, mine, d);
            b |= ((word >> lane) & 1u) << d;
        }
        const uint32_t x = (uint32_t)w * 32u + lane;
        if (x < g.dim_x) bytes[(size_t)r * g.dim_x + x] = (uint8_t)b;
    }
}

__global__ void __launch_bounds__(128) pack_cell_type_kernel(const int32_t* __restrict__ ct, uint32_t* __restrict__ ns,
                                                             uint32_t* __restrict__ sl, uint32_t* __restrict__ flags,
                                                             const Geom g, uint32_t row0)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.y * blockDim.x + threadIdx.x) >> 5;
    const uint32_t r = blockIdx.x;
    const uint32_t y = row0 + r;
    const int wpb = (gridDim.y * blockDim.x) >> 5;
    uint32_t any_ns = 0, any_sl = 0;
    for (int w = warp; w < (int)g.nw; w += wpb) {
        const uint32_t x = (uint32_t)w * 32u + lane;
        int t = 0;
        if (x < g.dim_x) t = ct[(size_t)r * g.dim_x + x];
        const uint32_t wn = __ballot_sync(0xFFFFFFFFu, t == 1);
        const uint32_t ws = __ballot_sync(0xFFFFFFFFu, t == 2);
        if (lane == 0) {
            ns[(size_t)y * g.pitch + w] = wn;
            sl[(size_t)y * g.pitch + w] = ws;
        }
        any_ns |= wn;
        any_sl |= ws;
    }
    if (lane == 0) {
        if (any_ns) atomicOr(&flags[0], 1u);
        if (any_sl) atomicOr(&flags[1], 1u);
    }
}

// chirality: bit (first_bit + r*dim_x + x) of the flat bit-field `bits` (which starts at bit `bit_base`
// of the uploaded chunk)
__global__ void __launch_bounds__(128) pack_rnd_kernel(const uint8_t* __restrict__ bits, uint32_t* __restrict__ ch,
                                                       const Geom g, uint32_t row0, uint64_t first_bit)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.y * blockDim.x + threadIdx.x) >> 5;
    const uint32_t r = blockIdx.x;
    const uint32_t y = row0 + r;
    const int wpb = (gridDim.y * blockDim.x) >> 5;
    for (int w = warp; w < (int)g.nw; w += wpb) {
        const uint32_t x = (uint32_t)w * 32u + lane;
        uint32_t bit = 0;
        if (x < g.dim_x) {
            const uint64_t i = first_bit + (uint64_t)r * g.dim_x + x;
            bit = (bits[i >> 3] >> (i & 7)) & 1u;
        }
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, bit);
        if (lane == 0) ch[(size_t)y * g.pitch + w] = word;
    }
}

// E/W domain-edge mask of one row: site 0 and site dim_x-1
__global__ void build_xedge_kernel(uint32_t* __restrict__ xedge, const Geom g)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= g.pitch) return;
    uint32_t m = 0;
    if (w == 0) m |= 1u;
    if (w == (g.dim_x - 1) / 32) m |= 1u << ((g.dim_x - 1) & 31);
    xedge[w] = m;
}

static dim3 pack_grid(const Geom& g, uint32_t nrows)
{
    uint32_t warps = (g.nw + 0) ;
    uint32_t blocks_y = (warps + 3) / 4;
    if (blocks_y > 64) blocks_y = 64;
    if (blocks_y < 1) blocks_y = 1;
    return dim3(nrows, blocks_y, 1);
}

int launch_pack_state(lgca_b200_lattice* h, const uint8_t* d_bytes, uint32_t* planes, uint32_t row0, uint32_t nrows,
                      cudaStream_t s)
{
    if (nrows == 0) return 0;
    dim3 grid = pack_grid(h->g, nrows);
    switch (h->nd) {
    case 4: pack_state_kernel<4><<<grid, 128, 0, s>>>(d_bytes, planes, h->g, row0); break;
    case 6: pack_state_kernel<6><<<grid, 128, 0, s>>>(d_bytes, planes, h->g, row0); break;
    default: pack_state_kernel<7><<<grid, 128, 0, s>>>(d_bytes, planes, h->g, row0); break;
    }
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_unpack_state(lgca_b200_lattice* h, const uint32_t* planes, uint8_t* d_bytes, uint32_t row0, uint32_t nrows,
                        cudaStream_t s)
{
    if (nrows == 0) return 0;
    dim3 grid = pack_grid(h->g, nrows);
    switch (h->nd) {
    case 4: unpack_state_kernel<4><<<grid, 128, 0, s>>>(planes, d_bytes, h->g, row0); break;
    case 6: unpack_state_kernel<6><<<grid, 128, 0, s>>>(planes, d_bytes, h->g, row0); break;
    default: unpack_state_kernel<7><<<grid, 128, 0, s>>>(planes, d_bytes, h->g, row0); break;
    }
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_pack_cell_type(lgca_b200_lattice* h, const int32_t* d_ct, uint32_t row0, uint32_t nrows, cudaStream_t s)
{
    if (nrows == 0) return 0;
    pack_cell_type_kernel<<<pack_grid(h->g, nrows), 128, 0, s>>>(d_ct, h->ns, h->sl, h->d_flags, h->g, row0);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_pack_rnd(lgca_b200_lattice* h, const uint8_t* d_bits, uint64_t first_bit, uint32_t row0, uint32_t nrows,
                    cudaStream_t s)
{
    if (nrows == 0) return 0;
    pack_rnd_kernel<<<pack_grid(h->g, nrows), 128, 0, s>>>(d_bits, h->ch, h->g, row0, first_bit);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_build_xedge(lgca_b200_lattice* h, cudaStream_t s)
{
    build_xedge_kernel<<<(h->g.pitch + 127) / 128, 128, 0, s>>>(h->xedge, h->g);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace lgca_b200
