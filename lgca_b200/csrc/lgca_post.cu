// Post-processing and diagnostics on bit-planes: per-cell fields, coarse-grained means (popcount
// reduction), mean velocity, particle count, and the gather/scatter helpers of the exact body force.
//
// Reference: OMP_Lattice<M>::cell_post_process  src/omp_lattice.cpp:360-394
//            OMP_Lattice<M>::mean_post_process  src/omp_lattice.cpp:397-454  (window: SURVEY.md A.6)
//            OMP_Lattice<M>::get_mean_velocity  src/omp_lattice.cpp:508-557
//            Lattice<M>::get_n_particles        src/lattice.cpp:180-195
//            OMP_Lattice<M>::apply_body_force   src/omp_lattice.cpp:254-346
#include "lgca_internal.h"

namespace lgca_b200 {

// float(sin(M_PI/3)) as the reference's `static constexpr Real SIN` (src/lgca_models.h:236)
#define LGCA_SIN_F 0.866025388f

// Lattice vectors, LATTICE_VEC_X/Y (src/lgca_models.h:50-51, :245-246, :452-453)
template <int ND> __device__ __forceinline__ float vec_x(int d)
{
    if (ND == 4) return d == 0 ? 1.0f : (d == 2 ? -1.0f : 0.0f);
    switch (d) { case 0: return 1.0f; case 1: return 0.5f; case 2: return -0.5f; case 3: return -1.0f;
                 case 4: return -0.5f; case 5: return 0.5f; default: return 0.0f; }
}
template <int ND> __device__ __forceinline__ float vec_y(int d)
{
    if (ND == 4) return d == 1 ? 1.0f : (d == 3 ? -1.0f : 0.0f);
    switch (d) { case 1: case 2: return LGCA_SIN_F; case 4: case 5: return -LGCA_SIN_F; default: return 0.0f; }
}

// ---- per-cell fields: one thread per site of the owned rows ---------------------------------------
template <int ND>
__global__ void __launch_bounds__(256) cell_fields_kernel(const uint32_t* __restrict__ planes, float* __restrict__ rho,
                                                          float* __restrict__ mom, const Geom g, uint32_t own_rows)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t r = blockIdx.y;
    if (x >= g.dim_x || r >= own_rows) return;
    const size_t   base = (size_t)(r + g.halo) * g.pitch + (x >> 5);
    const uint32_t bit  = x & 31;
    int   dens = 0;
    float mx = 0.0f, my = 0.0f;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const uint32_t ns = (__ldg(planes + (size_t)d * g.plane_stride + base) >> bit) & 1u;
        dens += (int)ns;
        // same operation order as the reference: product first, then a float accumulate
        mx = __fadd_rn(mx, __fmul_rn((float)ns, vec_x<ND>(d)));
        my = __fadd_rn(my, __fmul_rn((float)ns, vec_y<ND>(d)));
    }
    const size_t cell = (size_t)r * g.dim_x + x;
    if (rho) rho[cell] = (float)dens;
    if (mom) { mom[2 * cell] = mx; mom[2 * cell + 1] = my; }
}

// ---- coarse-grained means: one thread per coarse cell ----------------------------------------------
// Window of coarse cell (cx,cy), r = cg radius: anchor site (cx*2r, cy*2r); columns [0, r], rows
// [0, 2r] clipped at the top of the GLOBAL domain (the reference's `abs(pos_x_neighbor - pos_x) <= r`
// and `neighbor_idx < num_cells` tests, src/omp_lattice.cpp:423-436).  Density and momentum-x sums are
// exact integers / half-integers (popcounts); momentum-y is s * integer in the popcount path or the
// reference's sequential row-major float32 sum when EXACT.
template <int ND, bool EXACT>
__global__ void __launch_bounds__(128) mean_fields_kernel(const uint32_t* __restrict__ planes,
                                                          const uint32_t* __restrict__ ghost_row, float* __restrict__ mrho,
                                                          float* __restrict__ mmom, const Geom g, uint32_t cg,
                                                          uint32_t coarse_dim_x, uint32_t coarse_rows,
                                                          uint32_t coarse_row0)
{
    const uint32_t cx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cr = blockIdx.y;               // coarse row inside the strip
    if (cx >= coarse_dim_x || cr >= coarse_rows) return;
    const uint32_t cy = coarse_row0 + cr;         // global coarse row
    const uint32_t ax = cx * 2u * cg;
    const uint32_t ay = cy * 2u * cg;             // global anchor row
    const uint32_t x1 = ax + cg;                  // last window column (inclusive), < dim_x

    int   pc[7] = {0, 0, 0, 0, 0, 0, 0};
    int   nrows = 0;
    float my_seq = 0.0f;
    for (uint32_t dy = 0; dy <= 2u * cg; ++dy) {
        const uint32_t gy = ay + dy;
        if (gy >= g.dim_y) break;                 // neighbor_idx >= num_cells
        const uint32_t sy = gy - g.y0 + g.halo;   // stored row (strip rows + upper halo)
        ++nrows;
        // strips: the row just above the owned rows comes from the side copy taken at snapshot time ([plane][pitch])
        const bool   beyond = ghost_row != nullptr && sy >= g.rows - g.halo;
        const uint32_t* src = beyond ? ghost_row : planes;
        const size_t pstride = beyond ? (size_t)g.pitch : (size_t)g.plane_stride;
        const size_t rb = beyond ? 0 : (size_t)sy * g.pitch;
        for (uint32_t w = ax >> 5; w <= (x1 >> 5); ++w) {
            uint32_t m = 0xFFFFFFFFu;
            if (w == (ax >> 5)) m &= 0xFFFFFFFFu << (ax & 31);
            if (w == (x1 >> 5)) m &= low_mask((int)(x1 & 31) + 1);
            uint32_t v[7];
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                v[d] = __ldg(src + (size_t)d * pstride + rb + w) & m;
                pc[d] += __popc(v[d]);
            }
            if (EXACT && ND != 4) {
                // reference order: row-major over the window, per-cell value k*s with
                // k = n1 + n2 - n4 - n5 (exact in float32), accumulated sequentially
                uint32_t any = (v[1] | v[2] | v[4] | v[5]);
                while (any) {
                    const int b = __ffs(any) - 1;
                    any &= any - 1;
                    const int k = (int)((v[1] >> b) & 1u) + (int)((v[2] >> b) & 1u) - (int)((v[4] >> b) & 1u)
                                - (int)((v[5] >> b) & 1u);
                    my_seq = __fadd_rn(my_seq, __fmul_rn((float)k, LGCA_SIN_F));
                }
            }
        }
    }
    const float cnt = (float)(nrows * (int)(cg + 1));
    int dsum = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) dsum += pc[d];
    float mx, my;
    if (ND == 4) {
        mx = (float)(pc[0] - pc[2]);
        my = (float)(pc[1] - pc[3]);
    } else {
        mx = __fmul_rn((float)(2 * (pc[0] - pc[3]) + pc[1] + pc[5] - pc[2] - pc[4]), 0.5f);
        my = EXACT ? my_seq : __fmul_rn((float)(pc[1] + pc[2] - pc[4] - pc[5]), LGCA_SIN_F);
    }
    const size_t cc = (size_t)cr * coarse_dim_x + cx;
    if (mrho) mrho[cc] = __fdiv_rn((float)dsum, cnt);
    if (mmom) { mmom[2 * cc] = __fdiv_rn(mx, cnt); mmom[2 * cc + 1] = __fdiv_rn(my, cnt); }
}

// ---- coarse-grained means, popcount BLOCK reduction (the non-exact mode) ---------------------------------------------
// Block = 32 coarse cells of one coarse row x MF_SLICES row slices: a warp reads the same lattice row for 32 neighbouring
// windows (consecutive words: coalesced), the slices split the 2r+1 window rows, the per-plane popcounts meet in shared
// memory.  Same window, same integer sums, hence the same floats as mean_fields_kernel<ND, false>.
constexpr int MF_SLICES = 8;
template <int ND>
__global__ void __launch_bounds__(32 * MF_SLICES) mean_fields_block_kernel(const uint32_t* __restrict__ planes,
                                                                          const uint32_t* __restrict__ ghost_row,
                                                                          float* __restrict__ mrho, float* __restrict__ mmom,
                                                                          const Geom g, uint32_t cg, uint32_t coarse_dim_x,
                                                                          uint32_t coarse_row0)
{
    __shared__ int s_pc[MF_SLICES][ND][32];
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const uint32_t cx = blockIdx.x * 32 + tx, cr = blockIdx.y, cy = coarse_row0 + cr;
    const uint32_t ax = cx * 2u * cg, ay = cy * 2u * cg, x1 = ax + cg;
    int pc[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) pc[d] = 0;
    if (cx < coarse_dim_x) {
        for (uint32_t dy = ty; dy <= 2u * cg; dy += MF_SLICES) {
            const uint32_t gy = ay + dy;
            if (gy >= g.dim_y) break;
            const uint32_t sy = gy - g.y0 + g.halo;
            const bool     beyond = ghost_row != nullptr && sy >= g.rows - g.halo;
            const uint32_t* src = beyond ? ghost_row : planes;
            const size_t pstride = beyond ? (size_t)g.pitch : (size_t)g.plane_stride;
            const size_t rb = beyond ? 0 : (size_t)sy * g.pitch;
            for (uint32_t w = ax >> 5; w <= (x1 >> 5); ++w) {
                uint32_t m = 0xFFFFFFFFu;
                if (w == (ax >> 5)) m &= 0xFFFFFFFFu << (ax & 31);
                if (w == (x1 >> 5)) m &= low_mask((int)(x1 & 31) + 1);
#pragma unroll
                for (int d = 0; d < ND; ++d) pc[d] += __popc(__ldg(src + (size_t)d * pstride + rb + w) & m);
            }
        }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) s_pc[ty][d][tx] = pc[d];
    __syncthreads();
    if (ty != 0 || cx >= coarse_dim_x) return;
#pragma unroll
    for (int d = 0; d < ND; ++d)
        for (int sl = 1; sl < MF_SLICES; ++sl) pc[d] += s_pc[sl][d][tx];
    const uint32_t rows_in = min(2u * cg + 1u, g.dim_y - ay); // window rows inside the global domain
    const float cnt = (float)((int)rows_in * (int)(cg + 1));
    int dsum = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) dsum += pc[d];
    float mx, my;
    if (ND == 4) {
        mx = (float)(pc[0] - pc[2]);
        my = (float)(pc[1] - pc[3]);
    } else {
        mx = __fmul_rn((float)(2 * (pc[0] - pc[3]) + pc[1] + pc[5] - pc[2] - pc[4]), 0.5f);
        my = __fmul_rn((float)(pc[1] + pc[2] - pc[4] - pc[5]), LGCA_SIN_F);
    }
    const size_t cc = (size_t)cr * coarse_dim_x + cx;
    if (mrho) mrho[cc] = __fdiv_rn((float)dsum, cnt);
    if (mmom) { mmom[2 * cc] = __fdiv_rn(mx, cnt); mmom[2 * cc + 1] = __fdiv_rn(my, cnt); }
}

// ---- mean velocity (device reduction, double accumulation) -----------------------------------------
// out3 = { sum_x, sum_y, #fluid cells } over the owned rows.
template <int ND>
__global__ void __launch_bounds__(256) mean_velocity_kernel(const uint32_t* __restrict__ planes,
                                                            const uint32_t* __restrict__ ns,
                                                            const uint32_t* __restrict__ sl, double* __restrict__ out3,
                                                            const Geom g, uint32_t own_rows)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    double sx = 0.0, sy = 0.0, cnt = 0.0;
    if (x < g.dim_x) {
        for (uint32_t r = blockIdx.y; r < own_rows; r += gridDim.y) {
            const size_t   base = (size_t)(r + g.halo) * g.pitch + (x >> 5);
            const uint32_t bit  = x & 31;
            const uint32_t solid = ((__ldg(ns + base) | __ldg(sl + base)) >> bit) & 1u;
            if (solid) continue;
            cnt += 1.0;
            int   dens = 0;
            float mx = 0.0f, my = 0.0f;
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const uint32_t b = (__ldg(planes + (size_t)d * g.plane_stride + base) >> bit) & 1u;
                dens += (int)b;
                mx = __fadd_rn(mx, __fmul_rn((float)b, vec_x<ND>(d)));
                my = __fadd_rn(my, __fmul_rn((float)b, vec_y<ND>(d)));
            }
            if (dens > 0) {
                sx += (double)__fdiv_rn(mx, (float)dens);
                sy += (double)__fdiv_rn(my, (float)dens);
            }
        }
    }
    // block reduction
    __shared__ double sh[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx  += __shfl_down_sync(0xFFFFFFFFu, sx, o);
        sy  += __shfl_down_sync(0xFFFFFFFFu, sy, o);
        cnt += __shfl_down_sync(0xFFFFFFFFu, cnt, o);
    }
    if (lane == 0) { sh[0][warp] = sx; sh[1][warp] = sy; sh[2][warp] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += sh[0][i]; b += sh[1][i]; c += sh[2][i]; }
        atomicAdd(out3 + 0, a);
        atomicAdd(out3 + 1, b);
        atomicAdd(out3 + 2, c);
    }
}

// ---- particle count: popcount over the owned rows of all planes -------------------------------------
__global__ void __launch_bounds__(256) count_kernel(const uint32_t* __restrict__ planes, unsigned long long* out,
                                                    const Geom g, int nd, uint32_t own_rows)
{
    unsigned long long c = 0;
    const size_t per_plane = (size_t)own_rows * g.pitch;
    const size_t total     = per_plane * nd;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t d = i / per_plane, o = i % per_plane;
        c += __popc(__ldg(planes + d * g.plane_stride + (size_t)g.halo * g.pitch + o));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// ---- body force helpers -------------------------------------------------------------------------------
// gather: byte state of the drawn cells (bit 7 set when the cell is not FLUID or lies outside the strip)
template <int ND>
__global__ void gather_cells_kernel(const uint32_t* __restrict__ planes, const uint32_t* __restrict__ ns,
                                    const uint32_t* __restrict__ sl, const int32_t* __restrict__ cells, size_t n,
                                    uint8_t* __restrict__ out, const Geom g, uint32_t own_rows)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t cell = (uint32_t)cells[i];
    const uint32_t gy = cell / g.dim_x, x = cell % g.dim_x;
    if (gy < g.y0 || gy >= g.y0 + own_rows) { out[i] = 0x80; return; }
    const size_t   base = (size_t)(gy - g.y0 + g.halo) * g.pitch + (x >> 5);
    const uint32_t bit  = x & 31;
    uint32_t b = 0;
#pragma unroll
    for (int d = 0; d < ND; ++d) b |= ((__ldg(planes + (size_t)d * g.plane_stride + base) >> bit) & 1u) << d;
    if (((__ldg(ns + base) | __ldg(sl + base)) >> bit) & 1u) b |= 0x80u;
    out[i] = (uint8_t)b;
}

// scatter: cells[i] takes the new byte state bytes[i] (each cell at most once per call)
template <int ND>
__global__ void apply_cells_kernel(uint32_t* __restrict__ planes, const int32_t* __restrict__ cells,
                                   const uint8_t* __restrict__ bytes, size_t n, const Geom g, uint32_t own_rows)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t cell = (uint32_t)cells[i];
    const uint32_t gy = cell / g.dim_x, x = cell % g.dim_x;
    if (gy < g.y0 || gy >= g.y0 + own_rows) return;
    const size_t   base = (size_t)(gy - g.y0 + g.halo) * g.pitch + (x >> 5);
    const uint32_t bit  = x & 31;
    const uint32_t nb   = bytes[i];
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        uint32_t* w = planes + (size_t)d * g.plane_stride + base;
        const uint32_t old = (*w >> bit) & 1u;
        if (old != ((nb >> d) & 1u)) atomicXor(w, 1u << bit);
    }
}

#define DISPATCH_ND(nd, CALL4, CALL6, CALL7)                                                          \
    do { if ((nd) == 4) { CALL4; } else if ((nd) == 6) { CALL6; } else { CALL7; } } while (0)

int launch_cell_fields(lgca_b200_lattice* h, const uint32_t* planes, float* d_rho, float* d_mom, cudaStream_t s)
{
    const Geom& g = h->g;
    const uint32_t own = g.rows - 2 * g.halo;
    dim3 grid((g.dim_x + 255) / 256, own, 1);
    DISPATCH_ND(h->nd, (cell_fields_kernel<4><<<grid, 256, 0, s>>>(planes, d_rho, d_mom, g, own)),
                (cell_fields_kernel<6><<<grid, 256, 0, s>>>(planes, d_rho, d_mom, g, own)),
                (cell_fields_kernel<7><<<grid, 256, 0, s>>>(planes, d_rho, d_mom, g, own)));
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_mean_fields(lgca_b200_lattice* h, const uint32_t* planes, const uint32_t* ghost_row, float* d_mrho, float* d_mmom,
                       int exact, cudaStream_t s)
{
    const Geom& g = h->g;
    const uint32_t cg = h->cfg.cg_radius;
    if (cg == 0) return set_error(LGCA_B200_EINVAL, "coarse fields requested but cg_radius == 0");
    const uint32_t own = g.rows - 2 * g.halo;
    const uint32_t cdx = g.dim_x / (2 * cg), crows = own / (2 * cg), crow0 = g.y0 / (2 * cg);
    if (cdx == 0 || crows == 0) return 0;
    dim3 grid((cdx + 127) / 128, crows, 1);
#define MF(ND, EX) mean_fields_kernel<ND, EX><<<grid, 128, 0, s>>>(planes, ghost_row, d_mrho, d_mmom, g, cg, cdx, crows, crow0)
    if (exact) {
        DISPATCH_ND(h->nd, (MF(4, true)), (MF(6, true)), (MF(7, true)));
    } else {
        // popcount block reduction (32 coarse cells x MF_SLICES row slices per block)
        dim3 bgrid((cdx + 31) / 32, crows, 1);
#define MFB(ND) mean_fields_block_kernel<ND><<<bgrid, 32 * MF_SLICES, 0, s>>>(planes, ghost_row, d_mrho, d_mmom, g, cg, cdx, crow0)
        DISPATCH_ND(h->nd, (MFB(4)), (MFB(6)), (MFB(7)));
#undef MFB
    }
#undef MF
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_mean_velocity(lgca_b200_lattice* h, const uint32_t* planes, double* d_out3, cudaStream_t s)
{
    const Geom& g = h->g;
    const uint32_t own = g.rows - 2 * g.halo;
    LGCA_CUDA_CHECK(cudaMemsetAsync(d_out3, 0, 3 * sizeof(double), s));
    dim3 grid((g.dim_x + 255) / 256, own < 512 ? own : 512, 1);
    DISPATCH_ND(h->nd, (mean_velocity_kernel<4><<<grid, 256, 0, s>>>(planes, h->ns, h->sl, d_out3, g, own)),
                (mean_velocity_kernel<6><<<grid, 256, 0, s>>>(planes, h->ns, h->sl, d_out3, g, own)),
                (mean_velocity_kernel<7><<<grid, 256, 0, s>>>(planes, h->ns, h->sl, d_out3, g, own)));
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_count_particles(lgca_b200_lattice* h, const uint32_t* planes, unsigned long long* d_out, cudaStream_t s)
{
    const Geom& g = h->g;
    const uint32_t own = g.rows - 2 * g.halo;
    LGCA_CUDA_CHECK(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), s));
    count_kernel<<<(h->sm_count > 0 ? h->sm_count : 148) * 8, 256, 0, s>>>(planes, d_out, g, h->nd, own);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_gather_cells(lgca_b200_lattice* h, const uint32_t* planes, const int32_t* d_cells, size_t n, uint8_t* d_bytes,
                        cudaStream_t s)
{
    if (n == 0) return 0;
    const Geom& g = h->g;
    const uint32_t own = g.rows - 2 * g.halo;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    DISPATCH_ND(h->nd, (gather_cells_kernel<4><<<blocks, 256, 0, s>>>(planes, h->ns, h->sl, d_cells, n, d_bytes, g, own)),
                (gather_cells_kernel<6><<<blocks, 256, 0, s>>>(planes, h->ns, h->sl, d_cells, n, d_bytes, g, own)),
                (gather_cells_kernel<7><<<blocks, 256, 0, s>>>(planes, h->ns, h->sl, d_cells, n, d_bytes, g, own)));
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_apply_flips(lgca_b200_lattice* h, uint32_t* planes, const int32_t* d_cells, size_t n, cudaStream_t s)
{
    // d_cells holds n cell indices followed (at byte offset n*4) by n new byte states
    if (n == 0) return 0;
    const Geom& g = h->g;
    const uint32_t own = g.rows - 2 * g.halo;
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(d_cells + n);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    DISPATCH_ND(h->nd, (apply_cells_kernel<4><<<blocks, 256, 0, s>>>(planes, d_cells, bytes, n, g, own)),
                (apply_cells_kernel<6><<<blocks, 256, 0, s>>>(planes, d_cells, bytes, n, g, own)),
                (apply_cells_kernel<7><<<blocks, 256, 0, s>>>(planes, d_cells, bytes, n, g, own)));
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace lgca_b200
