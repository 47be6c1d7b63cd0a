// collide_and_propagate as a register-resident wavefront: K time steps per pass over HBM.
//
// Reference semantics: OMP_Lattice<M>::collide_and_propagate, src/omp_lattice.cpp:100-249 (periodic pull
// streaming per SURVEY.md A.2, then collide / bounce at the destination cell), applied K times.
//
// Design (B200-first, nothing like the reference's per-cell loop):
//   * One warp owns a band of 32 consecutive 32-site words (1024 sites) and marches down a chunk of
//     rows.  Each lane keeps, for every fused time level, the few words of the previous two rows that
//     the hexagonal stencil still needs (7 words per level for FHP, 4 for HPP) -- a time-skewed
//     wavefront, so a row read from HBM is pushed through K updates before it is written back.
//   * Streaming in x is a 1-bit funnel shift (SHF) whose carry bit comes from the neighbour lane by
//     warp shuffle; the odd/even row offset of the hexagonal lattice decides which planes shift, and is
//     resolved at compile time by unrolling the row loop by two.
//   * Collision + walls are the LOP3 networks of lgca_collide.cuh.
//   * Lanes 0 and 31 are halo lanes: every step invalidates one more bit at the band edges, so for
//     K <= 32 the 30 interior words stay exact.  Chunks overlap by K rows at both ends (recomputed).
//   * Periodic wrap: rows by index arithmetic; in x the edge lanes assemble their word from the
//     periodic images (fetch_word), which also covers widths that are not multiples of 32.
#include "lgca_internal.h"

namespace lgca_b200 {

constexpr int WAVE_VALID = 30; // interior lanes per warp

__device__ __forceinline__ uint32_t up1(uint32_t w)   // site x <- site x-1
{
    return __funnelshift_l(__shfl_up_sync(0xFFFFFFFFu, w, 1), w, 1);
}
__device__ __forceinline__ uint32_t down1(uint32_t w) // site x <- site x+1
{
    return __funnelshift_r(w, __shfl_down_sync(0xFFFFFFFFu, w, 1), 1);
}

struct WaveParams {
    int bands;          // bands per row
    int chunk_rows;     // output rows per chunk (even)
    int chunks;         // chunks per lattice
    int tiles;          // bands * chunks
};

template <int MODEL, int K, bool HAS_NS, bool HAS_SL>
__global__ void __launch_bounds__(128) step_wave_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                        const uint32_t* __restrict__ ns_p,
                                                        const uint32_t* __restrict__ sl_p,
                                                        const uint32_t* __restrict__ ch_p,
                                                        const uint32_t* __restrict__ xedge, const Geom g,
                                                        const WaveParams wp)
{
    constexpr int  ND  = num_dir_of(MODEL);
    constexpr bool HPP = rule_of(MODEL) == MODEL_HPP;
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= wp.tiles) return;
    const int band  = tile % wp.bands;
    const int chunk = tile / wp.bands;
    const int wi    = band * WAVE_VALID - 1 + lane;          // word column of this lane (may be -1 / >= nw)
    const int ya    = chunk * wp.chunk_rows;                 // first output row (even)
    const int yb    = min(ya + wp.chunk_rows, (int)g.rows);  // one past the last output row
    const int total = (yb - ya) + 2 * K;                     // level-0 rows to push through

    // plain in-row word or periodic image?
    const bool regular = (wi >= 0) && (wi < (int)g.nw - (g.rem ? 1 : 0));
    const bool store_lane = (lane >= 1) && (lane <= WAVE_VALID) && (wi < (int)g.nw);
    const uint32_t vmask = store_lane ? valid_mask(g, wi) : 0u;
    const uint32_t ew = HAS_SL ? fetch_word(xedge, wi, g) : 0u;

    auto load = [&](const uint32_t* __restrict__ row) -> uint32_t {
        return regular ? __ldg(row + wi) : fetch_word(row, wi, g);
    };
    const int rows = (int)g.rows;
    // level-0 row index modulo the stored rows, advanced incrementally (no division in the loop)
    int r0m = ya - K;
    if (r0m < 0) r0m += rows;

    // delay lines per level transition (level s-1 -> s uses index s-1)
    uint32_t C0[K], C1[K], C2[K], C3[K], C6[K], D1[K], D2[K];
#pragma unroll
    for (int s = 0; s < K; ++s) { C0[s] = C1[s] = C2[s] = C3[s] = C6[s] = D1[s] = D2[s] = 0u; }

    // software prefetch of the next level-0 row
    uint32_t nxt[7];
    {
        const size_t r = (size_t)r0m * g.pitch;
#pragma unroll
        for (int d = 0; d < ND; ++d) nxt[d] = load(in + (size_t)d * g.plane_stride + r);
    }
    --r0m; // advanced to the arriving row at the top of every iteration

    for (int j = 0; j < total; j += 2) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
            const int jc = j + jj;
            if (jc < total) {
                uint32_t a[7];
#pragma unroll
                for (int d = 0; d < ND; ++d) a[d] = nxt[d];
                const int r0 = ya - K + jc; // level-0 row that just arrived (unwrapped)
                if (++r0m >= rows) r0m -= rows; // ... and its stored index
                if (jc + 1 < total) {
                    const int rn = (r0m + 1 >= rows) ? r0m + 1 - rows : r0m + 1;
                    const size_t r = (size_t)rn * g.pitch;
#pragma unroll
                    for (int d = 0; d < ND; ++d) nxt[d] = load(in + (size_t)d * g.plane_stride + r);
                }
#pragma unroll
                for (int s = 1; s <= K; ++s) {
                    // level s-1 row q = r0-(s-1) has arrived in a[]; produce level s row q-1 = r0-s
                    uint32_t n[7];
#pragma unroll
                    for (int d = 0; d < 7; ++d) n[d] = 0u;
                    if (jc >= 2 * s) {
                        int ym = r0m - s;          // stored index of row r0 - s (needs rows >= K)
                        if (ym < 0) ym += rows;
                        // ya is even and stored-row parity equals global parity (halo is even)
                        const bool odd = ((K + jj + s) & 1) != 0;
                        if (HPP) {
                            n[0] = up1(C0[s - 1]);
                            n[2] = down1(C2[s - 1]);
                            n[1] = D1[s - 1];       // plane 1 of row y-1
                            n[3] = a[3];            // plane 3 of row y+1
                        } else {
                            n[0] = up1(C0[s - 1]);
                            n[3] = down1(C3[s - 1]);
                            if (ND == 7) n[6] = C6[s - 1];
                            if (!odd) {
                                n[1] = up1(D1[s - 1]);
                                n[2] = D2[s - 1];
                                n[4] = a[4];
                                n[5] = up1(a[5]);
                            } else {
                                n[1] = D1[s - 1];
                                n[2] = down1(D2[s - 1]);
                                n[4] = down1(a[4]);
                                n[5] = a[5];
                            }
                        }
                        const size_t rm = (size_t)ym * g.pitch;
                        const uint32_t p  = HPP ? 0u : load(ch_p + rm);
                        const uint32_t ns = HAS_NS ? load(ns_p + rm) : 0u;
                        const uint32_t sl = HAS_SL ? load(sl_p + rm) : 0u;
                        const uint32_t ns_row =
                            (HAS_SL && ((uint32_t)ym == g.row_south || (uint32_t)ym == g.row_north)) ? 0xFFFFFFFFu : 0u;
                        collide_and_walls<MODEL, HAS_NS, HAS_SL>(n, p, ns, sl, ew, ns_row);
                    }
                    // rotate the delay line of this level transition
                    if (HPP) {
                        D1[s - 1] = C1[s - 1];      // plane 1: row q-1 -> becomes row y-1 next time
                        C1[s - 1] = a[1];
                        C0[s - 1] = a[0];
                        C2[s - 1] = a[2];
                    } else {
                        D1[s - 1] = C1[s - 1];
                        D2[s - 1] = C2[s - 1];
                        C1[s - 1] = a[1];
                        C2[s - 1] = a[2];
                        C0[s - 1] = a[0];
                        C3[s - 1] = a[3];
                        if (ND == 7) C6[s - 1] = a[6];
                    }
#pragma unroll
                    for (int d = 0; d < ND; ++d) a[d] = n[d];
                }
                if (jc >= 2 * K && store_lane) {
                    const size_t ro = (size_t)(r0 - K) * g.pitch + wi;
#pragma unroll
                    for (int d = 0; d < ND; ++d) out[(size_t)d * g.plane_stride + ro] = a[d] & vmask;
                }
            }
        }
    }
}

static WaveParams plan(const lgca_b200_lattice* h, int k)
{
    const Geom& g = h->g;
    WaveParams wp;
    wp.bands = ((int)g.nw + WAVE_VALID - 1) / WAVE_VALID;
    // enough tiles to fill 148 SMs with ~16 warps each, but chunks long enough to amortise the 2K-row overlap
    const int target_tiles = 148 * 16;
    int chunks = (target_tiles + wp.bands - 1) / wp.bands;
    int rows = (int)g.rows;
    int cr = (rows + chunks - 1) / chunks;
    const int min_rows = 16 * k;
    if (cr < min_rows) cr = min_rows;
    static int env_cr = -1;
    if (env_cr < 0) { const char* e = getenv("LGCA_B200_CHUNK_ROWS"); env_cr = e ? atoi(e) : 0; }
    if (env_cr > 0) cr = env_cr;
    cr = (cr + 1) & ~1;
    if (cr > rows) cr = (rows + 1) & ~1;
    wp.chunk_rows = cr;
    wp.chunks = (rows + cr - 1) / cr;
    wp.tiles = wp.bands * wp.chunks;
    return wp;
}

bool wave_supported(const lgca_b200_lattice* h, int k)
{
    if (k < 1 || k > 4) return false;
    // strips keep an even halo so that stored-row parity equals global parity
    if ((h->g.halo & 1u) || (h->g.y0 & 1u)) return false;
    if (h->g.rows < 8 || (int)h->g.rows < 2 * k) return false;
    if (!h->g.wrap_y && (uint32_t)k > h->g.halo) return false;
    return true;
}

template <int MODEL, int K>
static int launch_mk(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, cudaStream_t s)
{
    const WaveParams wp = plan(h, K);
    const int warps_per_block = 4;
    dim3 block(32 * warps_per_block, 1, 1);
    dim3 grid((wp.tiles + warps_per_block - 1) / warps_per_block, 1, 1);
#define GO(NS, SL)                                                                                              \
    step_wave_kernel<MODEL, K, NS, SL><<<grid, block, 0, s>>>(in, out, h->ns, h->sl, h->ch, h->xedge, h->g, wp)
    if (h->has_sl) { if (h->has_ns) GO(true, true); else GO(false, true); }
    else           { if (h->has_ns) GO(true, false); else GO(false, false); }
#undef GO
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int MODEL>
static int launch_m(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s)
{
    switch (k) {
    case 1: return launch_mk<MODEL, 1>(h, in, out, s);
    case 2: return launch_mk<MODEL, 2>(h, in, out, s);
    case 3: return launch_mk<MODEL, 3>(h, in, out, s);
    case 4: return launch_mk<MODEL, 4>(h, in, out, s);
    }
    return set_error(LGCA_B200_EINVAL, "unsupported k_fuse %d", k);
}

int launch_step_wave(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s)
{
    switch (rule_of(h->cfg.model)) {
    case MODEL_HPP:   return launch_m<MODEL_HPP>(h, in, out, k, s);
    case MODEL_FHP_I: return launch_m<MODEL_FHP_I>(h, in, out, k, s);
    default:          return launch_m<MODEL_FHP_II>(h, in, out, k, s);
    }
}

} // namespace lgca_b200
