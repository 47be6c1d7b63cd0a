// collide_and_propagate, straightforward form: one thread per 32-site output word, one time step per
// pass, neighbour words fetched from global memory (L1/L2 absorb the re-reads).  Handles every lattice
// shape (any dim_x >= 1, odd widths, strips with halo rows) and every cell type; it is the generic
// path and the A/B partner of the wavefront kernel (lgca_step_wave.cu).
//
// Reference: OMP_Lattice<M>::collide_and_propagate, src/omp_lattice.cpp:100-249 -- periodic PULL
// streaming (offset tables src/lgca_models.h:292-365 in the closed row-wise form of SURVEY.md A.2)
// followed by collide / bounce at the destination cell.
#include "lgca_internal.h"

namespace lgca_b200 {

__device__ __forceinline__ uint32_t shift_up(uint32_t w, uint32_t left) { return __funnelshift_l(left, w, 1); }
__device__ __forceinline__ uint32_t shift_down(uint32_t w, uint32_t right) { return __funnelshift_r(w, right, 1); }

template <int MODEL, bool HAS_NS, bool HAS_SL>
__global__ void __launch_bounds__(128) step_simple_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                          const uint32_t* __restrict__ ns_p,
                                                          const uint32_t* __restrict__ sl_p,
                                                          const uint32_t* __restrict__ ch_p,
                                                          const uint32_t* __restrict__ xedge, const Geom g)
{
    constexpr int ND = num_dir_of(MODEL);
    const int wi = blockIdx.y * blockDim.x + threadIdx.x;
    const int y  = (int)g.halo + blockIdx.x; // owned rows only: ghost rows of a strip belong to the ring neighbours
    if (wi >= (int)g.nw) return;

    const uint32_t gy  = (g.y0 + g.dim_y - g.halo % g.dim_y + (uint32_t)y) % g.dim_y; // global row
    const bool     odd = gy & 1u;
    const size_t   rc  = (size_t)y * g.pitch;
    const size_t   rs  = (size_t)row_index(y - 1, g) * g.pitch; // southern neighbour row
    const size_t   rn  = (size_t)row_index(y + 1, g) * g.pitch; // northern neighbour row

    uint32_t n[7];
#define PL(d) (in + (size_t)(d) * g.plane_stride)
    if (rule_of(MODEL) == MODEL_HPP) {
        n[0] = shift_up(fetch_word(PL(0) + rc, wi, g), fetch_word(PL(0) + rc, wi - 1, g));
        n[2] = shift_down(fetch_word(PL(2) + rc, wi, g), fetch_word(PL(2) + rc, wi + 1, g));
        n[1] = fetch_word(PL(1) + rs, wi, g);
        n[3] = fetch_word(PL(3) + rn, wi, g);
    } else {
        n[0] = shift_up(fetch_word(PL(0) + rc, wi, g), fetch_word(PL(0) + rc, wi - 1, g));
        n[3] = shift_down(fetch_word(PL(3) + rc, wi, g), fetch_word(PL(3) + rc, wi + 1, g));
        if (!odd) {
            n[1] = shift_up(fetch_word(PL(1) + rs, wi, g), fetch_word(PL(1) + rs, wi - 1, g));
            n[2] = fetch_word(PL(2) + rs, wi, g);
            n[4] = fetch_word(PL(4) + rn, wi, g);
            n[5] = shift_up(fetch_word(PL(5) + rn, wi, g), fetch_word(PL(5) + rn, wi - 1, g));
        } else {
            n[1] = fetch_word(PL(1) + rs, wi, g);
            n[2] = shift_down(fetch_word(PL(2) + rs, wi, g), fetch_word(PL(2) + rs, wi + 1, g));
            n[4] = shift_down(fetch_word(PL(4) + rn, wi, g), fetch_word(PL(4) + rn, wi + 1, g));
            n[5] = fetch_word(PL(5) + rn, wi, g);
        }
        if (ND == 7) n[6] = fetch_word(PL(6) + rc, wi, g);
    }
#undef PL
    const uint32_t p      = (rule_of(MODEL) == MODEL_HPP) ? 0u : fetch_word(ch_p + rc, wi, g);
    const uint32_t ns     = HAS_NS ? fetch_word(ns_p + rc, wi, g) : 0u;
    const uint32_t sl     = HAS_SL ? fetch_word(sl_p + rc, wi, g) : 0u;
    const uint32_t ew     = HAS_SL ? fetch_word(xedge, wi, g) : 0u;
    const uint32_t ns_row = (HAS_SL && (gy == 0 || gy == g.dim_y - 1)) ? 0xFFFFFFFFu : 0u;

    collide_and_walls<MODEL, HAS_NS, HAS_SL>(n, p, ns, sl, ew, ns_row);

    const uint32_t vm = valid_mask(g, wi);
#pragma unroll
    for (int d = 0; d < ND; ++d) out[(size_t)d * g.plane_stride + rc + wi] = n[d] & vm;
}

template <int MODEL>
static int launch_model(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, cudaStream_t s)
{
    const Geom& g = h->g;
    dim3 block(128, 1, 1);
    dim3 grid(g.rows - 2 * g.halo, (g.nw + 127) / 128, 1);
#define GO(NS, SL)                                                                                              \
    step_simple_kernel<MODEL, NS, SL><<<grid, block, 0, s>>>(in, out, h->ns, h->sl, h->ch, h->xedge, g)
    if (h->has_sl) { if (h->has_ns) GO(true, true); else GO(false, true); }
    else           { if (h->has_ns) GO(true, false); else GO(false, false); }
#undef GO
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int MODEL>
static int prepare_model()
{
    cudaFuncAttributes a;
    LGCA_CUDA_CHECK(cudaFuncGetAttributes(&a, step_simple_kernel<MODEL, false, false>));
    LGCA_CUDA_CHECK(cudaFuncGetAttributes(&a, step_simple_kernel<MODEL, true, false>));
    LGCA_CUDA_CHECK(cudaFuncGetAttributes(&a, step_simple_kernel<MODEL, false, true>));
    LGCA_CUDA_CHECK(cudaFuncGetAttributes(&a, step_simple_kernel<MODEL, true, true>));
    return 0;
}

// forces the (lazily loaded) kernel images onto the device, see wave_prepare()
int simple_prepare(lgca_b200_lattice* h)
{
    switch (rule_of(h->cfg.model)) {
    case MODEL_HPP:    return prepare_model<MODEL_HPP>();
    case MODEL_FHP_I:  return prepare_model<MODEL_FHP_I>();
    default:           return prepare_model<MODEL_FHP_II>();
    }
}

int launch_step_simple(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, cudaStream_t s)
{
    switch (rule_of(h->cfg.model)) {
    case MODEL_HPP:    return launch_model<MODEL_HPP>(h, in, out, s);
    case MODEL_FHP_I:  return launch_model<MODEL_FHP_I>(h, in, out, s);
    default:           return launch_model<MODEL_FHP_II>(h, in, out, s);
    }
}

} // namespace lgca_b200
