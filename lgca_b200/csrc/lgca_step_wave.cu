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
//   * Periodic wrap: rows by index arithmetic; in x the edge lanes read the periodic image of their
//     word (resolved once per lane), which also covers widths that are not multiples of 32.
//   * The next level-0 row and the level-1 masks are prefetched one iteration ahead; the pipeline
//     fill (first 2K rows of a chunk) runs in a separate, predicated copy of the row body so that the
//     steady-state loop is branch-free.
#include <math.h>
#include <stdlib.h>

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

// Per-lane view of the periodic row: where this lane's 32 sites come from.
template <bool IRREG>
struct LaneSrc {
    int  wa, wb, sh, n1;
    bool regular;
    __device__ __forceinline__ uint32_t load(const uint32_t* __restrict__ row) const
    {
        uint32_t v = __ldg(row + wa);
        if (IRREG) {
            if (!regular) { // lanes at the row end of a width that is not a multiple of 32
                v = __funnelshift_r(v, __ldg(row + wb), sh);
                if (n1 < 32) v = (v & low_mask(n1)) | (__ldg(row) << n1);
            }
        }
        return v;
    }
};

// Everything one lane carries through the row loop.
template <int K>
struct WaveState {
    // delay lines of the level transition s-1 -> s (index s-1): planes of the previous row (C*) and
    // planes 1,2 of the row before that (D*)
    uint32_t C0[K], C1[K], C2[K], C3[K], C6[K], D1[K], D2[K];
    uint32_t nxt[7];          // prefetched level-0 row
    uint32_t m_p, m_ns, m_sl; // prefetched masks of the row that level 1 produces next
    int      r0m;             // stored index of the level-0 row that arrived last
};

// One iteration: level-0 row r0 arrives, every level s produces its row r0 - s, and the level-K row
// r0 - K is stored.  PAR = parity of the iteration index (compile time); WARM = pipeline still filling
// (levels whose inputs are not there yet are skipped, nothing is stored before iteration 2K).
template <int MODEL, int K, bool HAS_NS, bool HAS_SL, bool IRREG, int PAR, bool WARM>
__device__ __forceinline__ void wave_row(WaveState<K>& st, const LaneSrc<IRREG>& src, int jc, int ya,
                                         const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                         const uint32_t* __restrict__ ns_p, const uint32_t* __restrict__ sl_p,
                                         const uint32_t* __restrict__ ch_p, const Geom& g, uint32_t ew, bool store_lane,
                                         uint32_t vmask, int wi)
{
    constexpr int  ND  = num_dir_of(MODEL);
    constexpr bool HPP = rule_of(MODEL) == MODEL_HPP;
    const int rows = (int)g.rows;

    uint32_t a[7];
#pragma unroll
    for (int d = 0; d < ND; ++d) a[d] = st.nxt[d];
    // masks for level 1 (row r0 - 1) were prefetched during the previous iteration
    const uint32_t p1 = st.m_p, ns1 = st.m_ns, sl1 = st.m_sl;

    if (++st.r0m >= rows) st.r0m -= rows; // stored index of the arriving row r0
    {
        // prefetch the next level-0 row (r0 + 1) and the masks of row r0 (level 1 needs them next
        // time).  Past the end of the chunk this reads a valid but unneeded row.
        const int    rn = (st.r0m + 1 >= rows) ? st.r0m + 1 - rows : st.r0m + 1;
        const size_t r  = (size_t)rn * g.pitch;
#pragma unroll
        for (int d = 0; d < ND; ++d) st.nxt[d] = src.load(in + (size_t)d * g.plane_stride + r);
        const size_t rm = (size_t)st.r0m * g.pitch;
        if (!HPP) st.m_p = src.load(ch_p + rm);
        if (HAS_NS) st.m_ns = src.load(ns_p + rm);
        if (HAS_SL) st.m_sl = src.load(sl_p + rm);
    }
#pragma unroll
    for (int s = 1; s <= K; ++s) {
        // level s-1 row q = r0-(s-1) is in a[]; produce level s row y = q-1 = r0-s
        uint32_t n[7];
#pragma unroll
        for (int d = 0; d < 7; ++d) n[d] = 0u;
        if (!WARM || jc >= 2 * s) {
            // ya is even and stored-row parity equals global parity (halo and y0 are even)
            const bool odd = ((K + PAR + s) & 1) != 0;
            if (HPP) {
                n[0] = up1(st.C0[s - 1]);
                n[2] = down1(st.C2[s - 1]);
                n[1] = st.D1[s - 1];      // plane 1 of row y-1
                n[3] = a[3];              // plane 3 of row y+1
            } else {
                n[0] = up1(st.C0[s - 1]);
                n[3] = down1(st.C3[s - 1]);
                if (ND == 7) n[6] = st.C6[s - 1];
                if (!odd) {
                    n[1] = up1(st.D1[s - 1]);
                    n[2] = st.D2[s - 1];
                    n[4] = a[4];
                    n[5] = up1(a[5]);
                } else {
                    n[1] = st.D1[s - 1];
                    n[2] = down1(st.D2[s - 1]);
                    n[4] = down1(a[4]);
                    n[5] = a[5];
                }
            }
            int ym = st.r0m - s; // stored index of row r0 - s (rows >= 2K)
            if (ym < 0) ym += rows;
            uint32_t p = 0u, ns = 0u, sl = 0u;
            if (s == 1) {
                p = p1; ns = ns1; sl = sl1;
            } else {
                // deeper levels re-read their mask words (this lane loaded them s-1 iterations ago: L1 hits)
                const size_t rm = (size_t)ym * g.pitch;
                if (!HPP) p = src.load(ch_p + rm);
                if (HAS_NS) ns = src.load(ns_p + rm);
                if (HAS_SL) sl = src.load(sl_p + rm);
            }
            const uint32_t ns_row =
                (HAS_SL && ((uint32_t)ym == g.row_south || (uint32_t)ym == g.row_north)) ? 0xFFFFFFFFu : 0u;
            collide_and_walls<MODEL, HAS_NS, HAS_SL>(n, p, ns, sl, ew, ns_row);
        }
        // rotate the delay line of this level transition
        st.D1[s - 1] = st.C1[s - 1];
        st.C1[s - 1] = a[1];
        st.C0[s - 1] = a[0];
        if (HPP) {
            st.C2[s - 1] = a[2];
        } else {
            st.D2[s - 1] = st.C2[s - 1];
            st.C2[s - 1] = a[2];
            st.C3[s - 1] = a[3];
            if (ND == 7) st.C6[s - 1] = a[6];
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) a[d] = n[d];
    }
    if ((!WARM || jc >= 2 * K) && store_lane) {
        const size_t ro = (size_t)(ya - 2 * K + jc) * g.pitch + wi;
#pragma unroll
        for (int d = 0; d < ND; ++d) out[(size_t)d * g.plane_stride + ro] = IRREG ? (a[d] & vmask) : a[d];
    }
}

template <int MODEL, int K, bool HAS_NS, bool HAS_SL, bool IRREG>
__global__ void __launch_bounds__(128) step_wave_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                        const uint32_t* __restrict__ ns_p,
                                                        const uint32_t* __restrict__ sl_p,
                                                        const uint32_t* __restrict__ ch_p,
                                                        const uint32_t* __restrict__ xedge, const Geom g,
                                                        const WavePlan wp)
{
    constexpr int ND = num_dir_of(MODEL);
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= wp.tiles) return;
    const int band  = tile % wp.bands;
    const int chunk = tile / wp.bands;
    const int wi    = band * WAVE_VALID - 1 + lane;          // word column of this lane (may be -1 / >= nw)
    const int ya    = chunk * wp.chunk_rows;                 // first output row (even)
    const int yb    = min(ya + wp.chunk_rows, (int)g.rows);  // one past the last output row
    const int total = (yb - ya) + 2 * K;                     // level-0 rows to push through
    const int rows  = (int)g.rows;

    // Periodic images in x are resolved ONCE per lane:
    //   * width a multiple of 32: every lane reads one plain word at a wrapped index;
    //   * otherwise lanes whose 32 sites touch the row end assemble them from up to three words
    //     (two around bit position p, plus word 0 after the wrap) with precomputed indices/shifts.
    LaneSrc<IRREG> src;
    src.wa = wi; src.wb = 0; src.sh = 0; src.n1 = 32; src.regular = true;
    if (!IRREG) {
        if (src.wa < 0) src.wa += (int)g.nw;
        else if (src.wa >= (int)g.nw) src.wa %= (int)g.nw;
    } else {
        src.regular = (wi >= 0) && (wi < (int)g.nw - 1);
        if (!src.regular) {
            long long p = ((long long)wi * 32) % (long long)g.dim_x;
            if (p < 0) p += g.dim_x;
            src.wa = (int)(p >> 5);
            src.sh = (int)(p & 31);
            src.wb = min(src.wa + 1, (int)g.nw - 1);
            src.n1 = (int)min((long long)32, (long long)g.dim_x - p); // sites before the row end
        }
    }
    const bool     store_lane = (lane >= 1) && (lane <= WAVE_VALID) && (wi < (int)g.nw);
    const uint32_t vmask      = store_lane ? valid_mask(g, wi) : 0u;
    const uint32_t ew         = HAS_SL ? src.load(xedge) : 0u;

    WaveState<K> st;
#pragma unroll
    for (int s = 0; s < K; ++s) st.C0[s] = st.C1[s] = st.C2[s] = st.C3[s] = st.C6[s] = st.D1[s] = st.D2[s] = 0u;
    st.m_p = st.m_ns = st.m_sl = 0u;
    st.r0m = ya - K;
    if (st.r0m < 0) st.r0m += rows;
    {
        const size_t r = (size_t)st.r0m * g.pitch;
#pragma unroll
        for (int d = 0; d < ND; ++d) st.nxt[d] = src.load(in + (size_t)d * g.plane_stride + r);
#pragma unroll
        for (int d = ND; d < 7; ++d) st.nxt[d] = 0u;
    }
    --st.r0m; // wave_row advances it to the arriving row first thing

#define LGCA_ROW(PAR, WARM, JC)                                                                                      \
    wave_row<MODEL, K, HAS_NS, HAS_SL, IRREG, PAR, WARM>(st, src, JC, ya, in, out, ns_p, sl_p, ch_p, g, ew, store_lane, \
                                                        vmask, wi)
    // pipeline fill: iterations 0 .. 2K-1 (2K is even, so parities alternate from 0)
#pragma unroll 1
    for (int j = 0; j < 2 * K; j += 2) {
        LGCA_ROW(0, true, j);
        LGCA_ROW(1, true, j + 1);
    }
    // steady state
    int j = 2 * K;
#pragma unroll 1
    for (; j + 1 < total; j += 2) {
        LGCA_ROW(0, false, j);
        LGCA_ROW(1, false, j + 1);
    }
    if (j < total) LGCA_ROW(0, false, j); // odd number of rows (HPP lattices with odd height)
#undef LGCA_ROW
}

// Chunk height: enough tiles to fill the machine, long enough to amortise the K-row pipeline fill.
// Cost model per fused step: every tile runs (cr + k - 1) level-rows (the fill is trapezoidal); tiles run
// in rounds of `resident` warps per SM; an SM with fewer than `saturate` warps is latency-bound and
// modelled as proportionally slower.
static WavePlan make_plan(const lgca_b200_lattice* h, int k)
{
    const Geom& g = h->g;
    WavePlan wp;
    wp.bands = ((int)g.nw + WAVE_VALID - 1) / WAVE_VALID;
    const int rows = (int)g.rows;
    const char* e_cr = getenv("LGCA_B200_CHUNK_ROWS");
    const char* e_res = getenv("LGCA_B200_RESIDENT_WARPS");
    const double resident = e_res ? atof(e_res) : 16.0, saturate = 12.0;
    int    best_cr = (rows + 1) & ~1;
    double best = 1e300;
    for (int cr = 2 * k; cr <= rows + 1; cr += 2) {
        const int    chunks = (rows + cr - 1) / cr;
        const double w      = (double)chunks * wp.bands / 148.0; // warps per SM
        const double rounds = fmax(1.0, ceil(w / resident));
        const double eff    = fmin(1.0, (w / rounds) / saturate);
        const double cost   = (double)(cr + k - 1) * rounds / eff;
        if (cost <= best) { best = cost; best_cr = cr; }
    }
    int cr = best_cr;
    if (e_cr && atoi(e_cr) > 0) cr = (atoi(e_cr) + 1) & ~1;
    if (cr > rows) cr = (rows + 1) & ~1;
    wp.chunk_rows = cr;
    wp.chunks = (rows + cr - 1) / cr;
    wp.tiles = wp.bands * wp.chunks;
    return wp;
}

bool wave_supported(const lgca_b200_lattice* h, int k)
{
    if (k < 1 || k > LGCA_MAX_K) return false;
    // strips keep an even halo and start on an even row so that stored-row parity equals global parity
    if ((h->g.halo & 1u) || (h->g.y0 & 1u)) return false;
    if (h->g.rows < 8 || (int)h->g.rows < 2 * k) return false;
    if (h->g.dim_x < 64) return false; // tiny rows wrap more than once inside a word: generic kernel
    if (!h->g.wrap_y && (uint32_t)k > h->g.halo) return false;
    return true;
}

template <int MODEL, int K>
static int launch_mk(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, cudaStream_t s)
{
    if (!h->plan_valid[K]) { h->plans[K] = make_plan(h, K); h->plan_valid[K] = 1; }
    const WavePlan wp = h->plans[K];
    const int warps_per_block = 4;
    dim3 block(32 * warps_per_block, 1, 1);
    dim3 grid((wp.tiles + warps_per_block - 1) / warps_per_block, 1, 1);
    const bool irreg = h->g.rem != 0;
#define GO(NS, SL)                                                                                                \
    do {                                                                                                          \
        if (irreg) step_wave_kernel<MODEL, K, NS, SL, true><<<grid, block, 0, s>>>(in, out, h->ns, h->sl, h->ch,  \
                                                                                    h->xedge, h->g, wp);          \
        else step_wave_kernel<MODEL, K, NS, SL, false><<<grid, block, 0, s>>>(in, out, h->ns, h->sl, h->ch,       \
                                                                               h->xedge, h->g, wp);               \
    } while (0)
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
