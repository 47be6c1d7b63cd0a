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
//   * Level-0 rows are prefetched two iterations ahead, the level-1 masks one; the pipeline
//     fill (first 2K rows of a chunk) runs in a separate, predicated copy of the row body so that the
//     steady-state loop is branch-free.
//   * One warp per block on a 2-D grid (band, chunk): every row index and loop bound is warp-uniform and lives
//     on the uniform datapath; plane addresses are one IMAD.WIDE each (row pointer + plane stride * d).
//   * Wall variants dispatch, per tile, to the all-fluid tile body when the tile's whole input window holds no
//     solid site (flag map computed once per plan by tile_fluid_kernel).
//   * Tiling (make_plan): 2-4 rounds of resident warps per launch -- the scheduler serves its warps by strict
//     priority, so one long round ends with half-empty schedulers.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "lgca_internal.h"
#include "lgca_wave_pins.h"

namespace lgca_b200 {

constexpr int WAVE_VALID = 30; // interior lanes per warp

// x-streaming by one site.  The neighbour lane's word comes by warp shuffle, the 1-bit funnel shift is a SHF.
// Measured alternative (-DLGCA_FMA_SHIFT=1): two multiply-adds on the otherwise idle FMA pipe,
//     (w << 1) | (nb >> 31)  =  w * 2    + hi32(nb * 2)            IMAD + IMAD.HI
//     (w >> 1) | (nb << 31)  =  hi32(w * 2^31) + nb * 2^31          IMAD.HI + IMAD
// (the summands never share a bit, so + is |; the multipliers must be kernel ARGUMENTS, with literals ptxas
// strength-reduces the products back into SHF / LEA).  It takes 40 of 366 instructions per loop body off the
// integer pipe but adds 40 issue slots, and on B200 it is 3 % SLOWER (C5: 104.9 vs 101.5 us per update): the loop
// is bound by issue slots + integer pipe together, not by the integer pipe alone.
#ifndef LGCA_FMA_SHIFT
#define LGCA_FMA_SHIFT 0
#endif
__device__ __forceinline__ uint32_t up1(uint32_t w, uint32_t two)   // site x <- site x-1
{
    const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, w, 1);
#if LGCA_FMA_SHIFT
    return w * two + __umulhi(nb, two);
#else
    return __funnelshift_l(nb, w, 1);
#endif
}
__device__ __forceinline__ uint32_t down1(uint32_t w, uint32_t half) // site x <- site x+1
{
    const uint32_t nb = __shfl_down_sync(0xFFFFFFFFu, w, 1);
#if LGCA_FMA_SHIFT
    return __umulhi(w, half) + nb * half;
#else
    return __funnelshift_r(w, nb, 1);
#endif
}

// Kernel arguments: per-plane base pointers live in the constant parameter bank, so an address is one
// IMAD.WIDE (32-bit word offset * 4 + constant base) on the FMA pipe instead of integer-pipe LEA pairs.
struct WaveArgs {
    const uint32_t* in[7];
    uint32_t*       out[7];
    const uint32_t* ns;
    const uint32_t* sl;
    const uint32_t* ch;
    const uint32_t* xedge;
    // native ring: tiles that read ghost rows spin until the neighbours have published `ring_epoch`
    const uint32_t* ring_flags; // [0] from the lower neighbour, [1] from the upper, [4] my own last completed push; nullptr = no in-kernel wait
    uint32_t        ring_epoch;
    uint32_t        ring_push_epoch;   // edge tiles overwrite rows my push of this epoch read: wait until it has completed
    const uint8_t*  tile_fluid;        // wall variants: [chunks * bands] 1 = the tile's input window is all fluid (or nullptr)
    uint32_t        stride_bytes;      // distance between consecutive planes of a set (in[d] = in[0] + d * stride)
    uint32_t        mul_two, mul_half; // 2 and 2^31: run-time multipliers of the FMA-pipe funnel shifts (up1/down1)
    // chained launches (whole lattices): per-chunk completion counters of this plan.  A tile bumps its chunk's counter
    // when its rows are stored; a tile of the NEXT launch starts as soon as the chunks its input rows lie in have been
    // completed by all bands of this launch (`chain_target` = launches of this plan so far x bands).
    uint32_t*       chain_done;        // [chunks] or nullptr
    uint32_t        chain_target;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Plane loads.  Default: the read-only path (ld.global.nc) -- neighbouring bands share two words per row through L1.
// COH (the two EDGE chunks of a strip while the native ring is running): ghost rows are stored by the neighbour GPU
// during the kernel's lifetime, which the .nc path is not defined for (a narrow row shares its 128-byte line with rows
// an interior tile may have pulled into L1 before the push arrived), so those tiles load through L2 (ld.global.cg)
// after their acquire of the epoch flag.  Interior tiles never touch a ghost row (the prefetch is clamped to the
// tile's input rows) and keep the faster path.
// -DLGCA_WAVE_PLAIN_LD=1 (A-B builds): ordinary cached loads (ld.global.ca) instead of the read-only path everywhere.
#ifndef LGCA_WAVE_PLAIN_LD
#define LGCA_WAVE_PLAIN_LD 0
#endif
template <bool COH>
__device__ __forceinline__ uint32_t ld_plane(const uint32_t* p) { return COH ? __ldcg(p) : (LGCA_WAVE_PLAIN_LD ? __ldca(p) : __ldg(p)); }

// Per-lane view of the periodic row: where this lane's 32 sites come from.
template <bool IRREG, bool COH = false>
struct LaneSrc {
    uint32_t wa, wb;
    int      sh, n1;
    bool     regular;
    // row_off = word offset of the row start inside a plane
    __device__ __forceinline__ uint32_t load(const uint32_t* __restrict__ plane, uint32_t row_off) const
    {
        uint32_t v = ld_plane<COH>(plane + (row_off + wa));
        if (IRREG) {
            if (!regular) { // lanes at the row end of a width that is not a multiple of 32
                v = __funnelshift_r(v, ld_plane<COH>(plane + (row_off + wb)), sh);
                if (n1 < 32) v = (v & low_mask(n1)) | (ld_plane<COH>(plane + row_off) << n1);
            }
        }
        return v;
    }
    // All ND planes of one row.  The planes of a set are equally spaced (`stride_bytes` apart), so the row address is
    // formed ONCE (one constant-bank base + IMAD.WIDE) and every further plane is a single IMAD.WIDE
    // (stride * d + row pointer) -- no per-plane constant-bank pointer load.  Widths that are not a multiple of 32
    // keep the per-plane pointers (their edge lanes assemble a word from up to three loads).
    template <int ND>
    __device__ __forceinline__ void load_planes(uint32_t (&v)[7], const uint32_t* const (&planes)[7], uint32_t stride_bytes,
                                                uint32_t row_off) const
    {
        if (!IRREG) {
            const char* pr = (const char*)(planes[0] + (row_off + wa));
#pragma unroll
            for (int d = 0; d < ND; ++d) v[d] = ld_plane<COH>((const uint32_t*)(pr + (uint64_t)stride_bytes * (uint32_t)d));
        } else {
#pragma unroll
            for (int d = 0; d < ND; ++d) v[d] = load(planes[d], row_off);
        }
    }
};

// Everything one lane carries through the row loop.
template <int K>
struct WaveState {
    // delay lines of the level transition s-1 -> s (index s-1): planes of the previous row (C*) and
    // planes 1,2 of the row before that (D*)
    uint32_t C0[K], C1[K], C2[K], C3[K], C6[K], D1[K], D2[K];
    uint32_t nxt[7];                 // prefetched level-0 row r0 + 1
    uint32_t nx2[7];                 // prefetched level-0 row r0 + 2 (loads stay in flight for a whole iteration)
    uint32_t pm1, pns1, psl1;        // prefetched mask words of row r0 + 1
    uint32_t Mp[K], Mns[K], Msl[K];  // mask words of the rows the levels produce next: index s-1 <-> row r0 - s
    uint32_t r0m;                    // stored index of the level-0 row that arrived last (slip walls only)
    uint32_t ro;                     // word offset of the last level-0 row fetched
};

// One iteration: level-0 row r0 arrives, every level s produces its row r0 - s, and the level-K row
// r0 - K is stored.  PAR = parity of the iteration index (compile time); WARM = pipeline still filling
// (levels whose inputs are not there yet are skipped, nothing is stored before iteration 2K).
template <int MODEL, int K, bool HAS_NS, bool HAS_SL, bool IRREG, bool COH, int PAR, bool WARM>
__device__ __forceinline__ void wave_row(WaveState<K>& st, const LaneSrc<IRREG, COH>& src, int jc, int fetch_end,
                                         uint32_t plane_words, uint32_t out_off, const WaveArgs& A, const Geom& g,
                                         uint32_t ew, bool store_lane, uint32_t vmask)
{
    constexpr int  ND  = num_dir_of(MODEL);
    constexpr bool HPP = rule_of(MODEL) == MODEL_HPP;
    const uint32_t rows = g.rows;

    uint32_t a[7];
#pragma unroll
    for (int d = 0; d < ND; ++d) { a[d] = st.nxt[d]; st.nxt[d] = st.nx2[d]; }
    // mask words of the arriving row r0 (prefetched one iteration ago); level 1 uses them next iteration
    const uint32_t pm = st.pm1, pns = st.pns1, psl = st.psl1;

    if (HAS_SL) { if (++st.r0m >= rows) st.r0m -= rows; } // stored index of the arriving row r0 (slip walls only)
    // Prefetch TWO rows ahead: level-0 row r0 + 2 and the mask words of row r0 + 1.  The loads have a whole
    // iteration to land wherever the scheduler places them.
    {
        // the mask row r0 + 1 is the row the planes were prefetched from one iteration ago
        const uint32_t rm = st.ro;
        // never fetch past the tile's last input row (the last fetch is simply repeated): ghost rows of a strip
        // are only ever read by tiles that waited for them
        if (jc < fetch_end) { st.ro += g.pitch; if (st.ro >= plane_words) st.ro -= plane_words; }
        const uint32_t ro = st.ro;
        src.template load_planes<ND>(st.nx2, A.in, A.stride_bytes, ro);
        if (!HPP) st.pm1 = src.load(A.ch, rm);
        if (HAS_NS) st.pns1 = src.load(A.ns, rm);
        if (HAS_SL) st.psl1 = src.load(A.sl, rm);
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
                n[0] = up1(st.C0[s - 1], A.mul_two);
                n[2] = down1(st.C2[s - 1], A.mul_half);
                n[1] = st.D1[s - 1];      // plane 1 of row y-1
                n[3] = a[3];              // plane 3 of row y+1
            } else {
                n[0] = up1(st.C0[s - 1], A.mul_two);
                n[3] = down1(st.C3[s - 1], A.mul_half);
                if (ND == 7) n[6] = st.C6[s - 1];
                if (!odd) {
                    n[1] = up1(st.D1[s - 1], A.mul_two);
                    n[2] = st.D2[s - 1];
                    n[4] = a[4];
                    n[5] = up1(a[5], A.mul_two);
                } else {
                    n[1] = st.D1[s - 1];
                    n[2] = down1(st.D2[s - 1], A.mul_half);
                    n[4] = down1(a[4], A.mul_half);
                    n[5] = a[5];
                }
            }
            const uint32_t p  = HPP ? 0u : st.Mp[s - 1];
            const uint32_t ns = HAS_NS ? st.Mns[s - 1] : 0u;
            const uint32_t sl = HAS_SL ? st.Msl[s - 1] : 0u;
            if (HAS_NS || HAS_SL) {
                // walls are rare: skip their logic for warps whose 1024 sites are all fluid
                uint32_t in[7];
#pragma unroll
                for (int d = 0; d < 7; ++d) in[d] = n[d];
                collide<MODEL>(n, p);
                if (__any_sync(0xFFFFFFFFu, (ns | sl) != 0u)) {
                    uint32_t ns_row = 0u;
                    if (HAS_SL) {
                        uint32_t ym = st.r0m + rows - (uint32_t)s; // stored index of row r0 - s
                        if (ym >= rows) ym -= rows;
                        ns_row = (ym == g.row_south || ym == g.row_north) ? 0xFFFFFFFFu : 0u;
                    }
                    apply_walls<MODEL, HAS_NS, HAS_SL>(n, in, ns, sl, ew, ns_row);
                }
            } else {
                collide<MODEL>(n, p);
            }
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
    // masks travel down the levels with the rows they belong to
#pragma unroll
    for (int s = K - 1; s >= 1; --s) {
        if (!HPP) st.Mp[s] = st.Mp[s - 1];
        if (HAS_NS) st.Mns[s] = st.Mns[s - 1];
        if (HAS_SL) st.Msl[s] = st.Msl[s - 1];
    }
    st.Mp[0] = pm; st.Mns[0] = pns; st.Msl[0] = psl;

    if ((!WARM || jc >= 2 * K) && store_lane) {
        if (!IRREG) {
            char* po = (char*)(A.out[0] + out_off); // same addressing scheme as load_planes
#pragma unroll
            for (int d = 0; d < ND; ++d) *(uint32_t*)(po + (uint64_t)A.stride_bytes * (uint32_t)d) = a[d];
        } else {
#pragma unroll
            for (int d = 0; d < ND; ++d) A.out[d][out_off] = IRREG ? (a[d] & vmask) : a[d];
        }
    }
}

// Register cap via the resident-blocks hint.  One-warp blocks are spread over the four SM sub-partitions of 16384
// registers each, so the occupancy steps are 128 registers -> 4 warps per sub-partition, 96 -> 5, 80 -> 6, 72 -> 7, 64 -> 8.
// The all-fluid FHP variants (K <= 5) fit 96 registers with at most two spilled words, which lifts them from 4 to 5
// resident warps per scheduler (+7 % measured).  Measured and rejected: K = 6 at 96 registers (11 spilled words,
// -1.4 %), the wall variants at 96 (33-43 spilled words, -11 %), K = 4 at 80 registers (-4 %).
// The prologue / epilogue of chained launches (chain_wait / chain_signal, the ring's push wait) nudges ptxas over an
// occupancy step in about two dozen of the 160 variants, so every variant is pinned to the class it had without them
// (lgca_wave_pins.h, generated by scripts/gen_wave_pins.py from the ptxas log of the pre-chain build,
// profiles/r03_wave_ptxas_before_chain.log); variants the pin would make spill run one class lower instead.
template <int MODEL, int K, bool HAS_NS, bool HAS_SL, bool IRREG>
constexpr int wave_min_blocks()
{
    return wave_pin(rule_of(MODEL), K, HAS_NS, HAS_SL, IRREG);
}
// Row range [oa, ob) of a tile's output, relative to the first owned row.  Whole lattices: uniform chunks.
// Strips: tile rows 0 and 1 are the bottom and the top EDGE chunk (the only ones that read ghost rows); they are
// scheduled first and are shorter than the interior chunks, so that waiting for the neighbours' ghost rows at
// their start does not delay the end of the kernel.
__device__ __forceinline__ void tile_rows(const Geom& g, const WavePlan& wp, int c, int& oa, int& ob)
{
    const int owned = (int)(g.rows - 2 * g.halo);
    if (wp.edge_rows == 0) {
        oa = c * wp.chunk_rows;
        ob = min(oa + wp.chunk_rows, owned);
    } else if (c == 0) {
        oa = 0; ob = wp.edge_rows;
    } else if (c == 1) {
        oa = owned - wp.edge_rows; ob = owned;
    } else {
        oa = wp.edge_rows + (c - 2) * wp.chunk_rows;
        ob = min(oa + wp.chunk_rows, owned - wp.edge_rows);
    }
}

// Periodic images in x are resolved ONCE per lane:
//   * width a multiple of 32: every lane reads one plain word at a wrapped index;
//   * otherwise lanes whose 32 sites touch the row end assemble them from up to three words
//     (two around bit position p, plus word 0 after the wrap) with precomputed indices/shifts.
template <bool IRREG, bool COH = false>
__device__ __forceinline__ LaneSrc<IRREG, COH> make_lane_src(const Geom& g, int wi)
{
    LaneSrc<IRREG, COH> src;
    src.wb = 0; src.sh = 0; src.n1 = 32; src.regular = true;
    int wa = wi;
    if (!IRREG) {
        if (wa < 0) wa += (int)g.nw;
        else if (wa >= (int)g.nw) wa %= (int)g.nw;
    } else {
        src.regular = (wi >= 0) && (wi < (int)g.nw - 1);
        if (!src.regular) {
            long long p = ((long long)wi * 32) % (long long)g.dim_x;
            if (p < 0) p += g.dim_x;
            wa      = (int)(p >> 5);
            src.sh  = (int)(p & 31);
            src.wb  = (uint32_t)min(wa + 1, (int)g.nw - 1);
            src.n1  = (int)min((long long)32, (long long)g.dim_x - p); // sites before the row end
        }
    }
    src.wa = (uint32_t)wa;
    return src;
}

// Chained launches.  Consecutive launches of one plan on one stream are launched with programmatic stream
// serialisation and every block releases its dependents first thing, so the tiles of launch n+1 are dispatched into
// the warp slots the tiles of launch n leave behind -- the staggered end of a launch (3.1: the scheduler serves its
// warps by strict priority) is filled with the next launch instead of running on half-empty schedulers.  What orders
// the data is the per-chunk counter: tile (n+1, c) reads rows [oa - K, ob + K) of the planes launch n writes, i.e.
// rows of the chunks c-1 .. c+1 (periodic), and starts once all bands of those chunks have bumped their counters
// (release, per lane: stores -> __threadfence -> red.add; acquire, per lane: relaxed poll -> __threadfence, which also
// drops this SM's L1 lines; the counters count lanes, 32 per tile).  The same wait covers the write-after-read side: tile (n+1, c) writes
// rows of the buffer launch n READ, and the only launch-n tiles that read those rows are the ones it waited for.
// Rows a tile reads are not written again before it has bumped its own counter (their next writer is a tile of launch
// n+2 in the same neighbourhood), so they are constant for the tile's lifetime.  No deadlock: the dependents of a
// launch are dispatched only after ALL its blocks have started, so whatever a spinning tile waits for is resident or done.
__device__ __forceinline__ void chain_wait(const WaveArgs& A, const Geom& g, const WavePlan& wp, int c, int K)
{
    // every lane polls (same address: one broadcast request) -- no divergent region in front of the tile body
    int oa, ob;
    tile_rows(g, wp, c, oa, ob);
    const int owned = (int)(g.rows - 2 * g.halo);
    const int E     = wp.edge_rows;
    int n  = ob - oa + 2 * K;                          // input rows, from row oa - K on
    int ra = oa - K;
    if (E == 0) {                                      // whole lattice: periodic
        n = min(n, owned);
        if (ra < 0) ra += owned;
    } else {                                           // strip: ghost rows are covered by the neighbours' epoch flags
        if (ra < 0) { n += ra; ra = 0; }
        n = min(n, owned - ra);
    }
    uint32_t ns = 64;
    while (n > 0) {
        int cc, end;                                   // chunk of row ra and its last row + 1 (tile_rows inverted)
        if (E == 0)               { cc = ra / wp.chunk_rows; end = min((cc + 1) * wp.chunk_rows, owned); }
        else if (ra < E)          { cc = 0; end = E; }
        else if (ra >= owned - E) { cc = 1; end = owned; }
        else                      { cc = 2 + (ra - E) / wp.chunk_rows; end = min(E + (cc - 1) * wp.chunk_rows, owned - E); }
        // relaxed polls with back-off (a waiting tile must not cost the running ones anything); the fence below makes
        // the successful poll an acquire
        while ((int32_t)(ld_relaxed_gpu(A.chain_done + cc) - A.chain_target) < 0) {
            __nanosleep(ns);
            if (ns < 1024) ns *= 2;
        }
        n -= end - ra;
        ra = (E == 0 && end >= owned) ? 0 : end;
    }
    __threadfence();
}
__device__ __forceinline__ void chain_signal(const WaveArgs& A, int c)
{
    // every lane publishes its own stores (counters count lanes: 32 per tile)
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(A.chain_done + c) : "memory");
}

// One tile (band x chunk) of the fused-step kernel.
template <int MODEL, int K, bool HAS_NS, bool HAS_SL, bool IRREG, bool COH>
__device__ __forceinline__ void wave_tile(const WaveArgs& A, const Geom& g, const WavePlan& wp, const int band, const int c)
{
    constexpr int ND = num_dir_of(MODEL);
    const int lane = threadIdx.x;
    int oa, ob;
    if (A.chain_target != 0u) chain_wait(A, g, wp, c, K);
    tile_rows(g, wp, c, oa, ob);
    if (COH) {
        // in-kernel halo wait: only the edge tiles depend on the neighbours' pushes; everyone else starts at once
        if (lane == 0) {
            if (c == 0) while (ld_acquire_sys(A.ring_flags + 0) < A.ring_epoch) __nanosleep(64);
            else        while (ld_acquire_sys(A.ring_flags + 1) < A.ring_epoch) __nanosleep(64);
            // (WAR) the rows this tile stores were read by my own ghost-row push two blocks ago (chained launches: the
            // stream does not order this launch behind that push any more)
            while ((int32_t)(ld_acquire_sys(A.ring_flags + 4) - A.ring_push_epoch) < 0) __nanosleep(64);
        }
        __syncwarp();
    }
    const int wi    = band * WAVE_VALID - 1 + lane;          // word column of this lane (may be -1 / >= nw)
    // output rows = owned rows only: ghost rows of a strip are never written by the step kernel (the ring
    // neighbours store into them); halo and all chunk heights are even, so ya stays even
    const int ya    = (int)g.halo + oa;                      // first output row (even)
    const int yb    = (int)g.halo + ob;                      // one past the last output row
    const int total = (yb - ya) + 2 * K;                     // level-0 rows to push through
    const int rows  = (int)g.rows;

    const LaneSrc<IRREG, COH> src = make_lane_src<IRREG, COH>(g, wi);
    const bool     store_lane = (lane >= 1) && (lane <= WAVE_VALID) && (wi < (int)g.nw);
    const uint32_t vmask      = store_lane ? valid_mask(g, wi) : 0u;
    const uint32_t ew         = HAS_SL ? src.load(A.xedge, 0u) : 0u;

    WaveState<K> st;
#pragma unroll
    for (int s = 0; s < K; ++s) {
        st.C0[s] = st.C1[s] = st.C2[s] = st.C3[s] = st.C6[s] = st.D1[s] = st.D2[s] = 0u;
        st.Mp[s] = st.Mns[s] = st.Msl[s] = 0u;
    }
    int r0 = ya - K;
    if (r0 < 0) r0 += rows;
    {
        const uint32_t ro = (uint32_t)r0 * g.pitch;
        const uint32_t r1 = (uint32_t)(r0 + 1 >= rows ? r0 + 1 - rows : r0 + 1);
        const uint32_t ro1 = r1 * g.pitch;
#pragma unroll
        for (int d = 0; d < 7; ++d) {
            st.nxt[d] = d < ND ? src.load(A.in[d], ro) : 0u;
            st.nx2[d] = d < ND ? src.load(A.in[d], ro1) : 0u;
        }
        st.pm1  = rule_of(MODEL) != MODEL_HPP ? src.load(A.ch, ro) : 0u;
        st.pns1 = HAS_NS ? src.load(A.ns, ro) : 0u;
        st.psl1 = HAS_SL ? src.load(A.sl, ro) : 0u;
    }
    st.r0m = (uint32_t)(r0 == 0 ? rows - 1 : r0 - 1); // wave_row advances it to the arriving row first thing
    st.ro = (uint32_t)(r0 + 1 >= rows ? r0 + 1 - rows : r0 + 1) * g.pitch;
    const int      fetch_end   = total - 2;             // iterations 0 .. total-3 fetch a new row (r0 + 2)
    const uint32_t plane_words = g.rows * g.pitch;

    // word offset of the row stored by the current iteration (row ya - 2K + jc), advanced per row
    uint32_t out_off = (uint32_t)ya * g.pitch + (uint32_t)max(wi, 0);
#define LGCA_ROW(PAR, WARM, JC)                                                                                      \
    wave_row<MODEL, K, HAS_NS, HAS_SL, IRREG, COH, PAR, WARM>(st, src, JC, fetch_end, plane_words, out_off, A, g, ew, store_lane, vmask)
    // pipeline fill: iterations 0 .. 2K-1 store nothing (2K is even, so parities alternate from 0)
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
        out_off += g.pitch;
        LGCA_ROW(1, false, j + 1);
        out_off += g.pitch;
    }
    if (j < total) LGCA_ROW(0, false, j); // odd number of rows (HPP lattices with odd height)
#undef LGCA_ROW
    if (A.chain_done) chain_signal(A, blockIdx.y);
}

// The kernel: one warp per block (the tile index and with it every loop bound is provably warp-uniform, so the
// shuffles in the row loop compile to plain SHFL without convergence bookkeeping); 2-D grid (x = band, y = chunk;
// blocks are issued x-fastest, so the chunk order is the schedule order) without a division, so every row index stays
// on the uniform datapath.
// Wall variants: walls are rare (C3: two wall rows and a cylinder, C4: a frame).  A tile whose input window -- its
// rows incl. the K-row aprons, its 32 word columns incl. the halo lanes -- holds no solid site at all runs the
// ALL-FLUID tile body (no mask loads, no pre-collision copies, no votes, no wall muxes, no mask delay lines); the
// per-tile flags are computed once per plan by tile_fluid_kernel.  The choice is per block, so the two bodies never
// merge inside a loop.
template <int MODEL, int K, bool HAS_NS, bool HAS_SL, bool IRREG>
__global__ void __launch_bounds__(32, wave_min_blocks<MODEL, K, HAS_NS, HAS_SL, IRREG>()) step_wave_kernel(const WaveArgs A, const Geom g, const WavePlan wp)
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); // the next chained launch may fill freed warp slots
    const int band = blockIdx.x;
    const int c    = blockIdx.y;
    if (wp.edge_rows && A.ring_flags && c < 2) // edge chunk of a strip on the native ring: flag wait + coherent loads
        wave_tile<MODEL, K, HAS_NS, HAS_SL, IRREG, true>(A, g, wp, band, c);
    else if ((HAS_NS || HAS_SL) && A.tile_fluid != nullptr && A.tile_fluid[c * wp.bands + band] != 0)
        wave_tile<MODEL, K, false, false, IRREG, false>(A, g, wp, band, c);
    else
        wave_tile<MODEL, K, HAS_NS, HAS_SL, IRREG, false>(A, g, wp, band, c);
}

// tile_fluid[c * bands + band] = 1 when no solid site lies in the tile's input window (same geometry as wave_tile).
// (MODEL only makes the instantiation unique to the translation unit that launches it)
template <int MODEL, bool IRREG>
__global__ void __launch_bounds__(32) tile_fluid_kernel(const uint32_t* __restrict__ ns, const uint32_t* __restrict__ sl,
                                                        const Geom g, const WavePlan wp, const int K, uint8_t* __restrict__ out)
{
    const int band = blockIdx.x, c = blockIdx.y, lane = threadIdx.x;
    int oa, ob;
    tile_rows(g, wp, c, oa, ob);
    const int wi   = band * WAVE_VALID - 1 + lane;
    const int rows = (int)g.rows;
    const LaneSrc<IRREG> src = make_lane_src<IRREG>(g, wi);
    const int total = (ob - oa) + 2 * K;
    int r = (int)g.halo + oa - K;
    if (r < 0) r += rows;
    uint32_t acc = 0u;
    for (int j = 0; j < total; ++j) {
        const uint32_t ro = (uint32_t)r * g.pitch;
        acc |= src.load(ns, ro) | src.load(sl, ro);
        if (++r >= rows) r -= rows;
    }
    const bool solid = __any_sync(0xFFFFFFFFu, acc != 0u);
    if (lane == 0) out[c * wp.bands + band] = solid ? 0 : 1;
}

// Chunk height: enough tiles to fill the machine, long enough to amortise the K-row pipeline fill.
// Cost model per fused step (fitted to chunk-height sweeps on B200, profiles/r01b_chunk_sweep.md):
//   * every tile runs (cr + k - 1) level-rows (the fill is trapezoidal) plus about three rows' worth of prologue;
//   * tiles run in rounds of `resident` warps per SM (from the occupancy of the kernel variant); an SM holding fewer
//     warps than that hides less latency and is modelled as sub-linearly slower;
//   * the warp scheduler serves its resident warps by strict priority, so the warps of a round do NOT progress in
//     lockstep: they finish staggered and the end of the launch runs with fewer and fewer warps per scheduler
//     (ncu: 2.96 of 4 warps active on average for a one-round launch).  That tail costs about 0.15 of one round
//     whatever the round is long, so two to four shorter rounds beat one long one (C5: -5.6 %, C4: -9 %).
static WavePlan make_plan(const lgca_b200_lattice* h, int k, int resident_warps)
{
    const Geom& g = h->g;
    WavePlan wp;
    wp.bands = ((int)g.nw + WAVE_VALID - 1) / WAVE_VALID;
    const int rows = (int)(g.rows - 2 * g.halo); // owned rows
    // tiling overrides for sweeps (scripts/chunk_sweep.py): results are identical for every tiling; only builds with
    // -DLGCA_B200_TUNING read the environment
#ifdef LGCA_B200_TUNING
    const char* e_cr = getenv("LGCA_B200_CHUNK_ROWS");
    const char* e_res = getenv("LGCA_B200_RESIDENT_WARPS");
#else
    const char* e_cr = nullptr;
    const char* e_res = nullptr;
#endif
    const double resident = e_res ? atof(e_res) : (double)resident_warps;
    // chained launches (whole lattices) hide most of the staggered end of a launch behind the next one (C5 with chaining:
    // 246-row chunks = 2 rounds 1.230e13, 164 rows = 3 rounds 1.226e13, 490 rows = 1 round 1.13e13)
    const double tail = !(h->cfg.flags & LGCA_B200_FLAG_NO_CHAIN) ? 0.10 : 0.15;
    int    best_cr = (rows + 1) & ~1;
    double best = 1e300;
    for (int cr = 2 * k; cr <= rows + 1; cr += 2) {
        const int    chunks = (rows + cr - 1) / cr;
        if (chunks + 2 > 65535) continue; // chunks ride in gridDim.y
        const double w      = (double)chunks * wp.bands / (double)(h->sm_count > 0 ? h->sm_count : 148); // warps per SM
        const double rounds = fmax(1.0, ceil(w / resident));
        const double eff    = pow(fmin(1.0, (w / rounds) / resident), 0.6);
        const double cost   = (double)(cr + k - 1 + 3) * (rounds + tail) / eff;
        if (cost <= best) { best = cost; best_cr = cr; }
    }
    int cr = best_cr;
    if (e_cr && atoi(e_cr) > 0) cr = (atoi(e_cr) + 1) & ~1;
    if (cr > rows) cr = (rows + 1) & ~1;
    wp.chunk_rows = cr;
    wp.edge_rows = 0;
    wp.chunks = (rows + cr - 1) / cr;
    if (g.halo) {
        // strips: two edge chunks of about 2/3 of an interior chunk (>= the rows the neighbours need and >= K + halo
        // so that no interior chunk reads ghost rows), interior chunks in between
        int ce = ((2 * cr / 3) + 1) & ~1;
        const int ce_min = (int)((g.halo + (uint32_t)k + 1) & ~1u);
        if (ce < ce_min) ce = ce_min;
        if (2 * ce + 2 * k <= rows) {
            wp.edge_rows = ce;
            const int inner = rows - 2 * ce;
            wp.chunks = 2 + (inner + cr - 1) / cr;
        }
    }
    wp.tiles = wp.bands * wp.chunks;
    return wp;
}

#if !defined(LGCA_WAVE_TU) || LGCA_WAVE_TU == 0
bool wave_supported(const lgca_b200_lattice* h, int k)
{
    if (k < 1 || k > (rule_of(h->cfg.model) == MODEL_HPP ? LGCA_MAX_K : LGCA_MAX_K_FHP)) return false;
    // strips keep an even halo and start on an even row so that stored-row parity equals global parity
    if ((h->g.halo & 1u) || (h->g.y0 & 1u)) return false;
    if (h->g.rows < 8 || (int)h->g.rows < 2 * k) return false;
    if (h->g.dim_x < 64) return false; // tiny rows wrap more than once inside a word: generic kernel
    if (!h->g.wrap_y && (uint32_t)k > h->g.halo) return false;
    return true;
}
#endif

template <int MODEL, int K, bool NS, bool SL, bool IRREG>
static int launch_variant(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, cudaStream_t s, bool chain)
{
    auto kernel = step_wave_kernel<MODEL, K, NS, SL, IRREG>;
    if (!h->plan_valid[K]) {
        // (the occupancy query also forces the lazily loaded kernel image onto the device)
        int blocks = 0;
        LGCA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, 32, 0));
        h->plans[K] = make_plan(h, K, blocks > 0 ? blocks : 16);
        if (NS || SL) {
            // per-tile "all fluid" flags of this plan (the masks are static; every mask writer invalidates the plans)
            const WavePlan& p = h->plans[K];
            const size_t need = (size_t)p.tiles;
            if (h->tile_fluid_cap[K] < need) {
                if (h->tile_fluid[K]) cudaFree(h->tile_fluid[K]);
                h->tile_fluid[K] = nullptr; h->tile_fluid_cap[K] = 0;
                LGCA_CUDA_CHECK(cudaMalloc((void**)&h->tile_fluid[K], need));
                h->tile_fluid_cap[K] = need;
            }
            if (p.chunks > 65535) return set_error(LGCA_B200_EINVAL, "chunk plan exceeds gridDim.y (%d chunks)", p.chunks);
            tile_fluid_kernel<MODEL, IRREG><<<dim3(p.bands, p.chunks, 1), dim3(32, 1, 1), 0, h->s_compute>>>(h->ns, h->sl, h->g, p, K,
                                                                                              h->tile_fluid[K]);
            h->launches++;
            LGCA_CUDA_CHECK(cudaGetLastError());
        }
        // chained launches: per-chunk completion counters of this plan (whole lattices with uniform chunks)
        // Chaining pays where one launch fills the machine (C3: +23 %, C5: +4 %).  On smaller lattices the waiting
        // tiles of the next launches sit in otherwise free warp slots next to the running ones and the launch-to-launch
        // latency of the counters exceeds the short tail they would hide (measured -15 % .. +20 %, lattice by lattice:
        // profiles/r03c_chain_sweep.log), so those keep the plain stream order.
        h->chain_launches[K] = 0;
        double min_fill = 0.9;
#ifdef LGCA_B200_TUNING
        if (getenv("LGCA_B200_CHAIN_MIN_FILL")) min_fill = atof(getenv("LGCA_B200_CHAIN_MIN_FILL"));
#endif
        const double slots = (double)(h->sm_count > 0 ? h->sm_count : 148) * (blocks > 0 ? blocks : 16);
        if ((h->g.halo == 0 || h->plans[K].edge_rows > 0) && !(h->cfg.flags & LGCA_B200_FLAG_NO_CHAIN) &&
            ((double)h->plans[K].tiles >= min_fill * slots || (h->cfg.flags & LGCA_B200_FLAG_FORCE_CHAIN))) {
            const size_t need = (size_t)h->plans[K].chunks;
            if (h->chain_cap[K] < need) {
                if (h->chain_done[K]) cudaFree(h->chain_done[K]);
                h->chain_done[K] = nullptr; h->chain_cap[K] = 0;
                LGCA_CUDA_CHECK(cudaMalloc((void**)&h->chain_done[K], need * sizeof(uint32_t)));
                h->chain_cap[K] = need;
            }
            // (stream-ordered behind every earlier launch of the old plan)
            LGCA_CUDA_CHECK(cudaMemsetAsync(h->chain_done[K], 0, need * sizeof(uint32_t), h->s_compute));
        } else if (h->chain_done[K]) {
            cudaFree(h->chain_done[K]);
            h->chain_done[K] = nullptr; h->chain_cap[K] = 0;
        }
        h->plan_valid[K] = 1;
    }
    const WavePlan wp = h->plans[K];
    WaveArgs A;
    for (int d = 0; d < 7; ++d) {
        const size_t off = (size_t)(d < h->nd ? d : 0) * h->g.plane_stride;
        A.in[d]  = in + off;
        A.out[d] = out + off;
    }
    A.ns = h->ns; A.sl = h->sl; A.ch = h->ch; A.xedge = h->xedge;
    A.ring_flags = h->ring_inkernel_epoch ? (const uint32_t*)h->ring_flags : nullptr;
    A.ring_epoch = h->ring_inkernel_epoch;
    A.ring_push_epoch = h->ring_inkernel_epoch ? h->ring_inkernel_epoch - 1u : 0u; // epoch = pushes issued; the one before the last
    A.mul_two = 2u; A.mul_half = 0x80000000u;
    A.stride_bytes = (uint32_t)(h->g.plane_stride * sizeof(uint32_t));
    A.tile_fluid = (NS || SL) ? h->tile_fluid[K] : nullptr;
    if (!in) { // prepare only (wave_prepare): plan + module load, no launch.  The flag kernel too: a plan may be
               // re-made (mask upload) after the ring has started, and a lazy module load behind a spinning kernel can hang
        cudaFuncAttributes fa;
        LGCA_CUDA_CHECK(cudaFuncGetAttributes(&fa, tile_fluid_kernel<MODEL, IRREG>));
        return 0;
    }
    if (wp.chunks > 65535) return set_error(LGCA_B200_EINVAL, "chunk plan exceeds gridDim.y (%d chunks)", wp.chunks);
    A.chain_done   = s == h->s_compute ? h->chain_done[K] : nullptr;
    A.chain_target = h->chain_launches[K] * (uint32_t)wp.bands * 32u; // wraps with the counters (compared as a signed difference)
    if (A.chain_done && chain) {
        // the op before this one on `s` is a launch of the same plan: start as its blocks retire (see chain_wait)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(wp.bands, wp.chunks, 1);
        cfg.blockDim = dim3(32, 1, 1);
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        LGCA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, A, h->g, wp));
    } else {
        kernel<<<dim3(wp.bands, wp.chunks, 1), dim3(32, 1, 1), 0, s>>>(A, h->g, wp);
    }
    if (A.chain_done) h->chain_launches[K]++;
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int MODEL, int K, bool IRREG>
static int launch_mk(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, cudaStream_t s, bool chain)
{
#define GO(NS, SL) launch_variant<MODEL, K, NS, SL, IRREG>(h, in, out, s, chain)
    if (h->has_sl) return h->has_ns ? GO(true, true) : GO(false, true);
    return h->has_ns ? GO(true, false) : GO(false, false);
#undef GO
}

template <int MODEL, bool IRREG>
static int launch_m(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain)
{
    switch (k) {
    case 1: return launch_mk<MODEL, 1, IRREG>(h, in, out, s, chain);
    case 2: return launch_mk<MODEL, 2, IRREG>(h, in, out, s, chain);
    case 3: return launch_mk<MODEL, 3, IRREG>(h, in, out, s, chain);
    case 4: return launch_mk<MODEL, 4, IRREG>(h, in, out, s, chain);
    case 5: return launch_mk<MODEL, 5, IRREG>(h, in, out, s, chain);
    case 6: return launch_mk<MODEL, 6, IRREG>(h, in, out, s, chain);
    case 7: if (MODEL == MODEL_HPP) return launch_mk<MODEL_HPP, 7, IRREG>(h, in, out, s, chain); break;
    case 8: if (MODEL == MODEL_HPP) return launch_mk<MODEL_HPP, 8, IRREG>(h, in, out, s, chain); break;
    }
    return set_error(LGCA_B200_EINVAL, "unsupported k_fuse %d", k);
}

// The kernel variants are spread over six translation units (collision rule x regular/irregular width) that the
// build compiles in parallel: -DLGCA_WAVE_TU=0..5 selects one, no define = everything in one unit.
int launch_wave_hpp_reg(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain);
int launch_wave_hpp_irr(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain);
int launch_wave_fhp1_reg(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain);
int launch_wave_fhp1_irr(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain);
int launch_wave_fhp2_reg(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain);
int launch_wave_fhp2_irr(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain);
#if !defined(LGCA_WAVE_TU) || LGCA_WAVE_TU == 0
int launch_wave_hpp_reg(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain) { return launch_m<MODEL_HPP, false>(h, in, out, k, s, chain); }
#endif
#if !defined(LGCA_WAVE_TU) || LGCA_WAVE_TU == 1
int launch_wave_hpp_irr(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain) { return launch_m<MODEL_HPP, true>(h, in, out, k, s, chain); }
#endif
#if !defined(LGCA_WAVE_TU) || LGCA_WAVE_TU == 2
int launch_wave_fhp1_reg(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain) { return launch_m<MODEL_FHP_I, false>(h, in, out, k, s, chain); }
#endif
#if !defined(LGCA_WAVE_TU) || LGCA_WAVE_TU == 3
int launch_wave_fhp1_irr(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain) { return launch_m<MODEL_FHP_I, true>(h, in, out, k, s, chain); }
#endif
#if !defined(LGCA_WAVE_TU) || LGCA_WAVE_TU == 4
int launch_wave_fhp2_reg(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain) { return launch_m<MODEL_FHP_II, false>(h, in, out, k, s, chain); }
#endif
#if !defined(LGCA_WAVE_TU) || LGCA_WAVE_TU == 5
int launch_wave_fhp2_irr(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain) { return launch_m<MODEL_FHP_II, true>(h, in, out, k, s, chain); }
#endif

#if !defined(LGCA_WAVE_TU) || LGCA_WAVE_TU == 0
int launch_step_wave(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain)
{
    if ((uint64_t)h->g.rows * h->g.pitch > 0xFFFFFFFFull || (uint64_t)h->g.plane_stride * sizeof(uint32_t) > 0xFFFFFFFFull)
        return set_error(LGCA_B200_EINVAL, "plane too large for 32-bit word offsets");
    const bool irreg = h->g.rem != 0;
    switch (rule_of(h->cfg.model)) {
    case MODEL_HPP:   return irreg ? launch_wave_hpp_irr(h, in, out, k, s, chain) : launch_wave_hpp_reg(h, in, out, k, s, chain);
    case MODEL_FHP_I: return irreg ? launch_wave_fhp1_irr(h, in, out, k, s, chain) : launch_wave_fhp1_reg(h, in, out, k, s, chain);
    default:          return irreg ? launch_wave_fhp2_irr(h, in, out, k, s, chain) : launch_wave_fhp2_reg(h, in, out, k, s, chain);
    }
}

// true when the plan for k has edge chunks, i.e. the in-kernel ghost-row wait of the native ring is available
bool wave_has_edge_chunks(lgca_b200_lattice* h, int k)
{
    if (!wave_supported(h, k)) return false;
    if (!h->plan_valid[k] && launch_step_wave(h, nullptr, nullptr, k, 0, false) != 0) return false;
    // The spinning edge tiles are scheduled first.  If they alone could fill the resident-block capacity of the device,
    // the push kernel of the previous block (which the NEIGHBOUR's edge tiles wait for) might never be dispatched and
    // the ring would hang: very wide strips fall back to the stream-ordered wait kernel.
    const int sms = h->sm_count > 0 ? h->sm_count : 148;
    if (2 * h->plans[k].bands > 4 * sms) return false;
    return h->plans[k].edge_rows > 0;
}

// Plans every K the handle can use and forces the kernel images to be loaded.  With CUDA's lazy module loading
// the first launch of a kernel may have to wait for the device to go idle -- fatal if another stream of the
// same process is spinning in the ring's wait kernel; the ring calls this before it starts.
int wave_prepare(lgca_b200_lattice* h)
{
    for (int k = 1; k <= h->k_fuse; ++k) {
        if (!wave_supported(h, k)) continue;
        const int rc = launch_step_wave(h, nullptr, nullptr, k, 0, false);
        if (rc) return rc;
    }
    return 0;
}
#endif // LGCA_WAVE_TU == 0

} // namespace lgca_b200
