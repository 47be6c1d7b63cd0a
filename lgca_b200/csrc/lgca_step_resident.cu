// collide_and_propagate for lattices that FIT ON CHIP: the whole lattice lives in the shared memory of the SMs for
// the n steps of a call -- one launch, one HBM/L2 read and one write per call, whatever n is.
//
// Reference semantics: OMP_Lattice<M>::collide_and_propagate, src/omp_lattice.cpp:100-249 (periodic pull streaming
// per SURVEY.md A.2, then collide / bounce at the destination cell), applied n times.  The reference's own app sizes
// (pipe 1400x700 / 1480x740, Karman 4400x2200: apps/*/…_viewer.h) and BASELINE config C2 (HPP 4096^2) are 1-17 M sites =
// 0.7-8.5 MB of bit-planes: far below the 33 MB of shared memory a B200 has, while the HBM-streaming wavefront kernel
// (lgca_step_wave.cu) is launch- and latency-bound on them (one warp per SM at 1400x700).
//
// Design:
//   * The rows are cut into one strip per CTA (<= one CTA per SM, cooperative launch: all CTAs are co-resident).  A CTA
//     stages its strip plus H ghost rows on both sides, all planes and the static masks, into shared memory with TMA
//     bulk copies (cp.async.bulk global->shared, completion on an mbarrier) and keeps it there.
//   * One time step = every thread pulls the words of one 32-site output word from the current buffer (neighbour words
//     for the 1-bit x-streaming come from shared memory too), runs the LOP3 collision / wall network of
//     lgca_collide.cuh and stores into the other buffer; one __syncthreads per step.  Ghost rows are recomputed
//     redundantly (trapezoid: after s steps the outermost s rows are stale), so K steps need no communication.
//   * Every K steps the CTAs exchange ghost rows through L2.  Only NEIGHBOURS synchronise -- no grid-wide barrier --
//     and the message carries its own arrival flag: every 32-site word travels as one 8-byte store {word, tag}
//     (single-copy atomic), tag = running block number, and the receiver polls the slot until the tag matches
//     ("LL" protocol: one L2 write + one L2 read of latency, no fence, no separate flag; measured 5 us -> ~1 us per
//     exchange against bulk store + release flag + acquire + bulk load).  The exchange area is double-buffered by block
//     parity: a CTA overwrites block b's slots with block b+2's rows only after it consumed its neighbours' block b+1
//     rows, which they sent after consuming block b (same argument as the multi-GPU ring, lgca_ring.cu).
//   * Row ends: words are row-aligned here, so widths that are not a multiple of 32 only change where the carry bit
//     of the x-shift comes from at the first / last word of a row.
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "lgca_internal.h"

namespace lgca_b200 {

#ifndef LGCA_RES_THREADS
#define LGCA_RES_THREADS 768
#endif
constexpr int RES_THREADS = LGCA_RES_THREADS;   // four words per thread and step (85 registers each); measured 512 / 768 / 1024: 768 is best or close on every lattice
constexpr int RES_MAX_SMEM = 227 * 1024; // opt-in dynamic shared memory per CTA on sm_100

struct ResArgs {
    const uint32_t* in;       // [nd][rows][pitch]
    uint32_t*       out;
    const uint32_t* ns;
    const uint32_t* sl;
    const uint32_t* ch;
    const uint32_t* xedge;
    uint2*          exch;     // [2 parities][G][2 sides][nd][H][nw] of {word, tag}
    uint32_t        epoch_base; // tags are monotonic over launches: tag of block b = epoch_base + b + 1
    uint32_t        rows, pitch, nw, rem, dim_y_south, dim_y_north; // rows of the lattice; stored rows of the N/S domain edges
    uint32_t        plane_stride;  // words between planes in global memory
    int             G;        // CTAs
    int             unit;     // strip heights are multiples of unit (2 for the hexagonal models)
    int             base_units, extra_units; // CTA j owns (base + (j < extra)) units
    int             H, K;     // ghost rows per side, steps per exchange
    int             rows_max; // rows of the shared-memory buffers = max strip height + 2H
    int             n_steps;
};

#ifdef LGCA_RES_TIMING
// development aid (scripts/build_variant.sh x.so -DLGCA_RES_TIMING): cycles of CTA j's thread 0 in {poll, steps, publish, all}
__device__ unsigned long long g_res_timing[160][4];
#define RES_T(var) const long long var = clock64()
#define RES_ACC(i, a, b) res_acc[i] += (unsigned long long)((b) - (a))
#define RES_DECL unsigned long long res_acc[4] = {0, 0, 0, 0}
#define RES_FLUSH do { if (tid == 0) for (int i = 0; i < 4; ++i) g_res_timing[j][i] += res_acc[i]; } while (0)
#else
#define RES_T(var) do { } while (0)
#define RES_ACC(i, a, b) do { } while (0)
#define RES_DECL do { } while (0)
#define RES_FLUSH do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy writes (st.shared / ld.acquire results) visible to the async proxy (TMA) and vice versa
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// one 8-byte message {word, tag}: stored and loaded as a single access (single-copy atomic), straight to / from L2
__device__ __forceinline__ void st_msg(uint2* p, uint32_t word, uint32_t tag)
{
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(word), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_msg(const uint2* p)
{
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

// first row and height of CTA j's strip
__device__ __host__ __forceinline__ void res_strip(const ResArgs& A, int j, int& y0, int& R)
{
    const int before = j * A.base_units + (j < A.extra_units ? j : A.extra_units);
    y0 = before * A.unit;
    R  = (A.base_units + (j < A.extra_units ? 1 : 0)) * A.unit;
}

// ---- four words (128 sites) of one row per thread -----------------------------------------------------------------
// Everything is addressed in groups of four consecutive words: one LDS.128 / STS.128 per plane, the index arithmetic
// is paid once per 128 sites.  Rows are `pitch` words long (a multiple of 4, padding words are zero); the last real
// word of a row (index nw-1, possibly partial: rem = dim_x % 32 sites) may sit anywhere inside the last group.
struct GroupCtx {
    uint32_t left_off;  // word offset (inside the row) of the word left of the group: 4g-1, or nw-1 for the first group
    uint32_t lsh;       // pre-shift that brings the last site of the row to bit 31 (first group of a partial row)
    uint32_t right_off; // word offset of the word right of the group (unused in the last group)
    bool     lastg;     // the group holds the last real word of the row
    int      il;        // its position inside the group
    uint32_t hi;        // bit of that word which receives site 0 in a down-shift (periodic wrap)
};

__device__ __forceinline__ uint4 lds4(const uint32_t* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void  sts4(uint32_t* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

// site x <- site x-1 by `sh` (0 or 1) sites: row = plane row base (smem), g4 = 4*g
__device__ __forceinline__ void shift_up4(uint32_t (&o)[4], const uint32_t* row, uint32_t g4, const GroupCtx& c, uint32_t sh)
{
    const uint4    v = lds4(row + g4);
    const uint32_t L = row[c.left_off] << c.lsh;
    o[0] = __funnelshift_l(L, v.x, sh);
    o[1] = __funnelshift_l(v.x, v.y, sh);
    o[2] = __funnelshift_l(v.y, v.z, sh);
    o[3] = __funnelshift_l(v.z, v.w, sh);
}
// site x <- site x+1 by `sh` (0 or 1) sites
__device__ __forceinline__ void shift_down4(uint32_t (&o)[4], const uint32_t* row, uint32_t g4, const GroupCtx& c, uint32_t sh)
{
    const uint4 v = lds4(row + g4);
    uint32_t Rr = 0u;
    if (!c.lastg) Rr = row[c.right_off];
    o[0] = __funnelshift_r(v.x, v.y, sh);
    o[1] = __funnelshift_r(v.y, v.z, sh);
    o[2] = __funnelshift_r(v.z, v.w, sh);
    o[3] = __funnelshift_r(v.w, Rr, sh);
    if (c.lastg) {
        // periodic wrap: the last real word takes site 0 of the row into its top site (the words after it are padding)
        const uint32_t wrap = (row[0] & sh) << c.hi;
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] |= (i == c.il) ? wrap : 0u;
    }
}
__device__ __forceinline__ void plain4(uint32_t (&o)[4], const uint32_t* row, uint32_t g4)
{
    const uint4 v = lds4(row + g4);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}

// OWN selects the thread mapping of a step:
//   OWN = 0: groups of four consecutive words (one LDS.128 / STS.128 per plane) handed out dynamically, e = tid, tid + T, ...;
//            every step pays the (row, group) index arithmetic again (~170 instructions per word).
//   OWN = S > 0: STATIC ownership -- thread t owns the words t, t + T, ... (at most S) of the CTA's local rows for the whole
//            launch.  Word offsets, row-end flags, the row's hexagonal parity and the static masks of an owned word are
//            computed / loaded ONCE and stay in registers; a step is a row-range test, the plane loads, the shifts, the
//            LOP3 network and the stores.  Measured with clock64 (scripts/res_timing.py): the dynamic mapping spends
//            ~1450 cycles per step on C1 (660 words per CTA: issue-bound on its own index arithmetic).
template <int OWN> struct ResThreads { static constexpr int value = OWN > 0 ? 1024 : RES_THREADS; };

template <int MODEL, bool HAS_NS, bool HAS_SL, int OWN>
__global__ void __launch_bounds__(ResThreads<OWN>::value, 1) step_resident_kernel(const ResArgs A)
{
    constexpr int RES_THREADS = ResThreads<OWN>::value; // shadows the default inside the kernel
    constexpr int  ND  = num_dir_of(MODEL);
    constexpr bool HPP = rule_of(MODEL) == MODEL_HPP;
    constexpr int  NM  = (HPP ? 0 : 1) + (HAS_NS ? 1 : 0) + (HAS_SL ? 1 : 0); // static mask planes held on chip
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j   = blockIdx.x;
    int y0, R;
    res_strip(A, j, y0, R);
    const int H = A.H, P = (int)A.pitch, nw = (int)A.nw, rem = (int)A.rem;
    const int LR = R + 2 * H;                               // local rows of this CTA
    const uint32_t plane_sz = (uint32_t)A.rows_max * P;     // words per plane in shared memory
    uint32_t* buf0 = smem;
    uint32_t* buf1 = smem + (size_t)ND * plane_sz;
    uint32_t* msk  = smem + (size_t)2 * ND * plane_sz;      // [NM][rows_max][P]: ch, ns, sl (those present)
    const uint32_t* m_ch = msk;
    const uint32_t* m_ns = msk + (size_t)(HPP ? 0 : 1) * plane_sz;
    const uint32_t* m_sl = msk + (size_t)((HPP ? 0 : 1) + (HAS_NS ? 1 : 0)) * plane_sz;

    uint32_t phase = 0;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- stage the strip + ghost rows (periodic in y) and the masks into shared memory: TMA bulk copies ----------
    if (tid == 0) {
        const uint32_t row_bytes = (uint32_t)P * 4u;
        mbar_expect_tx(&bar, (uint32_t)(ND + NM) * (uint32_t)LR * row_bytes);
        int gy = (y0 - H) % (int)A.rows;
        if (gy < 0) gy += (int)A.rows;
        int lr = 0;
        while (lr < LR) {
            const int run = min(LR - lr, (int)A.rows - gy);
            const size_t goff = (size_t)gy * P, soff = (size_t)lr * P;
#pragma unroll
            for (int d = 0; d < ND; ++d) bulk_g2s(buf0 + d * plane_sz + soff, A.in + (size_t)d * A.plane_stride + goff, run * row_bytes, &bar);
            int m = 0;
            if (!HPP) bulk_g2s(msk + (m++) * plane_sz + soff, A.ch + goff, run * row_bytes, &bar);
            if (HAS_NS) bulk_g2s(msk + (m++) * plane_sz + soff, A.ns + goff, run * row_bytes, &bar);
            if (HAS_SL) bulk_g2s(msk + (m++) * plane_sz + soff, A.sl + goff, run * row_bytes, &bar);
            lr += run;
            gy = 0;
        }
    }
    // the second buffer starts out zero (padding words of rows that are never computed travel with the rows)
    for (uint32_t e = tid; e < (uint32_t)ND * plane_sz / 4; e += RES_THREADS) sts4(buf1 + 4 * e, make_uint4(0u, 0u, 0u, 0u));
    mbar_wait(&bar, phase);
    phase ^= 1;
    __syncthreads();

    // thread -> (row, group) stepping: element index e = tid, tid + T, ...; (dq, dr) = divmod(T, groups per row)
    const int PG = P / 4, gl = (nw - 1) / 4;                 // groups per row; group that holds the last real word
    const int dq = RES_THREADS / PG, dr = RES_THREADS % PG;
    const int q0 = tid / PG, g0 = tid % PG;
    // static ownership: word offset inside a plane, local row (-1 = slot unused), flags {first word, last word, odd row,
    // N/S domain-edge row, pre-shift of the partial last word << 8} and the static masks of every owned word
    constexpr int OWN_N = OWN > 0 ? OWN : 1;
    uint32_t own_off[OWN_N], own_flg[OWN_N], own_ch[HPP ? 1 : OWN_N], own_ns[HAS_NS ? OWN_N : 1], own_sl[HAS_SL ? OWN_N : 1],
        own_ew[HAS_SL ? OWN_N : 1];
    int own_row[OWN_N];
    if constexpr (OWN > 0) {
#pragma unroll
        for (int i = 0; i < OWN; ++i) {
            const int e = tid + i * RES_THREADS;
            const int r = e / nw, w = e - r * nw;
            const bool valid = r >= 1 && r < LR - 1; // rows 0 and LR-1 are never recomputed (their neighbours lie outside)
            own_row[i] = valid ? r : -1;
            own_off[i] = (uint32_t)(r * P + w);
            uint32_t f = (w == 0 ? 1u : 0u) | (w == nw - 1 ? 2u : 0u) | (((uint32_t)r & 1u) << 2) | ((w == 0 && rem) ? (32u - (uint32_t)rem) << 8 : 0u);
            if (!HPP) own_ch[HPP ? 0 : i] = valid ? m_ch[own_off[i]] : 0u;
            if (HAS_NS) own_ns[HAS_NS ? i : 0] = valid ? m_ns[own_off[i]] : 0u;
            if (HAS_SL) {
                own_sl[HAS_SL ? i : 0] = valid ? m_sl[own_off[i]] : 0u;
                own_ew[HAS_SL ? i : 0] = valid ? __ldg(A.xedge + w) : 0u;
                int gy = y0 - H + r;
                if (gy < 0) gy += (int)A.rows; else if (gy >= (int)A.rows) gy -= (int)A.rows;
                if ((uint32_t)gy == A.dim_y_south || (uint32_t)gy == A.dim_y_north) f |= 8u;
            }
            own_flg[i] = f;
        }
    }
    const uint32_t hi_last = rem ? (uint32_t)rem - 1u : 31u;
    const uint32_t vm_last = rem ? ((1u << rem) - 1u) : 0xFFFFFFFFu;

    uint32_t* cur = buf0;
    uint32_t* nxt = buf1;
    const int n_blocks = (A.n_steps + A.K - 1) / A.K;
    const int lower = (j + A.G - 1) % A.G, upper = (j + 1) % A.G;
    // Exchange area: [block parity][CTA][side][plane][H rows x P words] of {word, tag}.  A segment (one plane of one side)
    // mirrors H consecutive rows of the shared-memory plane, padding words included, so a message index inside a segment
    // is also the word offset at the destination: no index arithmetic per message.  Warps are dealt round-robin to the
    // 2 * ND segments (warp -> segment and chunk are computed once), lanes run over the segment's messages.
    const int    seg_n    = H * P;                                  // messages of one segment
    const int    side_n   = ND * seg_n;                             // messages of one side of one CTA
    const size_t parity_n = (size_t)A.G * 2 * side_n;
    constexpr int NWARPS = RES_THREADS / 32;
    constexpr int NSEG   = 2 * ND;
    static_assert(NWARPS >= NSEG, "one warp per exchange segment at least");
    const int x_seg = warp % NSEG, x_chunk = warp / NSEG;           // my segment, my first 32-message chunk in it
    const int x_nchunk = NWARPS / NSEG + (x_seg < NWARPS % NSEG ? 1 : 0); // warps sharing my segment
    const int x_half = x_seg / ND, x_d = x_seg % ND;                // 0: bottom side / lower ghost rows, 1: top / upper
    const int x_stride = x_nchunk * 32, x_m0 = x_chunk * 32 + lane; // my messages of a segment: x_m0 + q * x_stride < seg_n
    const int x_nmine = x_m0 < seg_n ? (seg_n - x_m0 + x_stride - 1) / x_stride : 0;
    const int x_n0 = (seg_n - x_chunk * 32 + x_stride - 1) / x_stride; // ... of lane 0 (never fewer than any other lane's)

    RES_DECL;
    RES_T(t_all0);
    for (int b = 0; b < n_blocks; ++b) {
        const int kb = min(A.K, A.n_steps - b * A.K);
        RES_T(t_p0);
        if (b > 0) {
            // ghost rows of this block = the neighbours' edge rows after block b-1: poll every message until its tag
            // shows.  One warp per (side, plane, row): consecutive lanes read consecutive 8-byte messages.
            const uint32_t tag = A.epoch_base + (uint32_t)b;
            const uint2* ex = A.exch + (size_t)((b - 1) & 1) * parity_n;
            // lower ghost rows [0, H) <- the lower neighbour's TOP side; upper ghost rows [H+R, LR) <- the upper neighbour's BOTTOM side
            const uint2* src = ex + ((size_t)(x_half ? upper : lower) * 2 + (x_half ? 0 : 1)) * side_n + (size_t)x_d * seg_n;
            uint32_t*    dst = cur + x_d * plane_sz + (size_t)(x_half ? H + R : 0) * P;
            // Waiting: ONE lane per warp polls ONE sentinel message (the last one the sender's warp of the same index stored),
            // with a short back-off -- 148 x 1024 threads polling all their messages would ask L2 for more than it can
            // deliver and delay the very stores they wait for.  Then every thread loads its messages (four in flight) and
            // re-polls only the ones whose tag has not arrived yet (stores may overtake each other).  The loop is kept lean
            // on purpose: ncu showed an 8-way unrolled version of it issuing more instructions than the steps themselves.
            if (lane == 0 && x_n0 > 0) {
                const uint2* sentinel = src + x_m0 - lane + (x_n0 - 1) * x_stride;
                while (ld_msg(sentinel).y != tag) __nanosleep(32);
            }
            __syncwarp();
            const uint2* sp = src + x_m0;
            uint32_t*    dp = dst + x_m0;
            for (int q0 = 0; q0 < x_nmine; q0 += 4) {
                uint2 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[q] = make_uint2(0u, tag);
                    if (q0 + q < x_nmine) v[q] = ld_msg(sp + (q0 + q) * x_stride);
                }
                bool missing;
                do {
                    missing = false;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (v[q].y != tag) { v[q] = ld_msg(sp + (q0 + q) * x_stride); missing = true; }
                } while (missing);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (q0 + q < x_nmine) dp[(q0 + q) * x_stride] = v[q].x;
            }
            __syncthreads();
        }
        RES_T(t_p1);
        RES_ACC(0, t_p0, t_p1);
        for (int s = 1; s <= kb; ++s) {
            // rows that are still needed and still valid after step s of this block
            const int ra = H - (kb - s), rb = H + R + (kb - s);
            if constexpr (OWN > 0) {
#pragma unroll
                for (int i = 0; i < OWN; ++i) {
                    if (own_row[i] < ra || own_row[i] >= rb) continue;
                    const uint32_t f = own_flg[i], o = own_off[i];
                    const bool     firstw = f & 1u, lastw = f & 2u;
                    const uint32_t odd = (f >> 2) & 1u, even = odd ^ 1u, lsh = f >> 8;
                    const uint32_t ol = firstw ? o + (uint32_t)nw - 1u : o - 1u;  // word left / right of mine (periodic)
                    const uint32_t orr = lastw ? o - ((uint32_t)nw - 1u) : o + 1u;
                    // site x <- site x-1 / site x+1 by sh (0 or 1) sites; dr = row displacement in words
                    auto up = [&](const uint32_t* pl, int dr, uint32_t sh) { return __funnelshift_l(pl[(int)ol + dr] << lsh, pl[(int)o + dr], sh); };
                    auto down = [&](const uint32_t* pl, int dr, uint32_t sh) {
                        const uint32_t v = pl[(int)o + dr], nb = pl[(int)orr + dr];
                        return lastw ? ((v >> sh) | ((nb & sh) << hi_last)) : __funnelshift_r(v, nb, sh);
                    };
                    uint32_t n[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
                    if (HPP) {
                        n[0] = up(cur + 0 * plane_sz, 0, 1u);
                        n[2] = down(cur + 2 * plane_sz, 0, 1u);
                        n[1] = cur[1 * plane_sz + o - P];
                        n[3] = cur[3 * plane_sz + o + P];
                    } else {
                        n[0] = up(cur + 0 * plane_sz, 0, 1u);
                        n[3] = down(cur + 3 * plane_sz, 0, 1u);
                        n[1] = up(cur + 1 * plane_sz, -P, even);
                        n[2] = down(cur + 2 * plane_sz, -P, odd);
                        n[4] = down(cur + 4 * plane_sz, P, odd);
                        n[5] = up(cur + 5 * plane_sz, P, even);
                        if (ND == 7) n[6] = cur[6 * plane_sz + o];
                    }
                    collide_and_walls<MODEL, HAS_NS, HAS_SL>(n, HPP ? 0u : own_ch[HPP ? 0 : i], HAS_NS ? own_ns[HAS_NS ? i : 0] : 0u,
                                                             HAS_SL ? own_sl[HAS_SL ? i : 0] : 0u, HAS_SL ? own_ew[HAS_SL ? i : 0] : 0u,
                                                             (f & 8u) ? 0xFFFFFFFFu : 0u);
                    const uint32_t vm = lastw ? vm_last : 0xFFFFFFFFu;
#pragma unroll
                    for (int d = 0; d < ND; ++d) nxt[d * plane_sz + o] = n[d] & vm;
                }
            } else {
                const int total = (rb - ra) * PG;
                int r = ra + q0, g = g0;
                for (int e = tid; e < total; e += RES_THREADS) {
                    const uint32_t rc = (uint32_t)r * P, rs = rc - P, rn = rc + P, g4 = 4u * (uint32_t)g;
                    GroupCtx c;
                    c.lastg     = g >= gl;
                    c.il        = (nw - 1) - 4 * gl;
                    c.hi        = hi_last;
                    c.left_off  = g == 0 ? (uint32_t)nw - 1u : g4 - 1u;
                    c.lsh       = (g == 0 && rem) ? 32u - (uint32_t)rem : 0u;
                    c.right_off = g4 + 4u;
                    uint32_t in[7][4];
                    if (HPP) {
                        shift_up4(in[0], cur + 0 * plane_sz + rc, g4, c, 1u);
                        shift_down4(in[2], cur + 2 * plane_sz + rc, g4, c, 1u);
                        plain4(in[1], cur + 1 * plane_sz + rs, g4);
                        plain4(in[3], cur + 3 * plane_sz + rn, g4);
                    } else {
                        // odd/even hexagonal rows, branch-free: even rows shift planes 1 and 5 up, odd rows shift planes 2 and
                        // 4 down; local parity == global parity (strip starts and H are even); a funnel shift by 0 passes through
                        const uint32_t odd = (uint32_t)r & 1u, even = odd ^ 1u;
                        shift_up4(in[0], cur + 0 * plane_sz + rc, g4, c, 1u);
                        shift_down4(in[3], cur + 3 * plane_sz + rc, g4, c, 1u);
                        shift_up4(in[1], cur + 1 * plane_sz + rs, g4, c, even);
                        shift_down4(in[2], cur + 2 * plane_sz + rs, g4, c, odd);
                        shift_down4(in[4], cur + 4 * plane_sz + rn, g4, c, odd);
                        shift_up4(in[5], cur + 5 * plane_sz + rn, g4, c, even);
                        if (ND == 7) plain4(in[6], cur + 6 * plane_sz + rc, g4);
                    }
                    uint32_t pch[4] = {0u, 0u, 0u, 0u}, pns[4] = {0u, 0u, 0u, 0u}, psl[4] = {0u, 0u, 0u, 0u}, pew[4] = {0u, 0u, 0u, 0u};
                    if (!HPP) plain4(pch, m_ch + rc, g4);
                    if (HAS_NS) plain4(pns, m_ns + rc, g4);
                    uint32_t ns_row = 0u;
                    if (HAS_SL) {
                        plain4(psl, m_sl + rc, g4);
                        const uint4 ev = __ldg(reinterpret_cast<const uint4*>(A.xedge + g4));
                        pew[0] = ev.x; pew[1] = ev.y; pew[2] = ev.z; pew[3] = ev.w;
                        int gy = y0 - H + r;                              // global (stored) row of this local row
                        if (gy < 0) gy += (int)A.rows; else if (gy >= (int)A.rows) gy -= (int)A.rows;
                        ns_row = ((uint32_t)gy == A.dim_y_south || (uint32_t)gy == A.dim_y_north) ? 0xFFFFFFFFu : 0u;
                    }
                    uint32_t out[7][4];
    #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t n[7];
    #pragma unroll
                        for (int d = 0; d < 7; ++d) n[d] = d < ND ? in[d][i] : 0u;
                        collide_and_walls<MODEL, HAS_NS, HAS_SL>(n, pch[i], pns[i], psl[i], pew[i], ns_row);
                        // words after the last real word are padding (kept zero), the last real word may be partial
                        const uint32_t vm = !c.lastg ? 0xFFFFFFFFu : (i < c.il ? 0xFFFFFFFFu : (i == c.il ? vm_last : 0u));
    #pragma unroll
                        for (int d = 0; d < ND; ++d) out[d][i] = n[d] & vm;
                    }
    #pragma unroll
                    for (int d = 0; d < ND; ++d) sts4(nxt + d * plane_sz + rc + g4, make_uint4(out[d][0], out[d][1], out[d][2], out[d][3]));
                    r += dq; g += dr;
                    if (g >= PG) { g -= PG; ++r; }
                }
            }
            __syncthreads();
            uint32_t* t = cur; cur = nxt; nxt = t;
        }
        RES_T(t_p2);
        RES_ACC(1, t_p1, t_p2);
        if (b + 1 < n_blocks) {
            // publish my edge rows for the neighbours' next block (the last step ended with a __syncthreads)
            const uint32_t tag = A.epoch_base + (uint32_t)b + 1u;
            // side 0 = BOTTOM: my lowest H owned rows [H, 2H); side 1 = TOP: my highest H owned rows [R, R+H)
            uint2*          dst = A.exch + (size_t)(b & 1) * parity_n + ((size_t)j * 2 + x_half) * side_n + (size_t)x_d * seg_n;
            const uint32_t* src = cur + x_d * plane_sz + (size_t)(x_half ? R : H) * P;
            for (int q = 0; q < x_nmine; ++q) st_msg(dst + x_m0 + q * x_stride, src[x_m0 + q * x_stride], tag);
        }
        RES_T(t_p3);
        RES_ACC(2, t_p2, t_p3);
    }
    RES_T(t_all1);
    RES_ACC(3, t_all0, t_all1);
    RES_FLUSH;
    // ---- write the strip back: TMA bulk stores -------------------------------------------------------------------
    if (tid == 0) {
        fence_proxy_async(); // the st.shared of the last step -> visible to the bulk stores
#pragma unroll
        for (int d = 0; d < ND; ++d)
            bulk_s2g(A.out + (size_t)d * A.plane_stride + (size_t)y0 * P, cur + d * plane_sz + (size_t)H * P, (uint32_t)R * P * 4u);
        bulk_commit();
        bulk_wait_all();
    }
}

// ---- planning ----------------------------------------------------------------------------------------------------
struct ResPlan {
    int    ok;
    int    G, unit, base_units, extra_units, H, K, rows_max, own;
    size_t smem_bytes, exch_words;
};

static ResPlan res_plan(const lgca_b200_lattice* h)
{
    ResPlan p;
    memset(&p, 0, sizeof(p));
    const Geom& g = h->g;
    if (g.halo != 0 || !g.wrap_y) return p;                       // whole lattices only
    if (h->cfg.flags & (LGCA_B200_FLAG_SIMPLE_KERNEL | LGCA_B200_FLAG_NO_RESIDENT)) return p;
    // Used wherever the lattice fits on chip (measured on B200, profiles/r02*_small_lattices.md: from the 256^2 app
    // default up to HPP 4096^2 = C2 and FHP-II 4096x2048 it beats the HBM-streaming wavefront kernel, by 1.1x at the
    // top end and 5x on the reference's pipe).  LGCA_B200_FLAG_NO_RESIDENT / FORCE_RESIDENT: A-B tests.
    const bool hpp = rule_of(h->cfg.model) == MODEL_HPP;
    const int  nd = h->nd, nm = (hpp ? 0 : 1) + (h->has_ns ? 1 : 0) + (h->has_sl ? 1 : 0);
    const int  unit = hpp ? 1 : 2;
    const int  rows = (int)g.rows, sms = h->sm_count > 0 ? h->sm_count : 148;
    if (rows % unit) return p;
    // Steps per ghost-row exchange K (H = K ghost rows per side, even for the hexagonal models): cfg.k_fuse when given,
    // else the K in {8, 6, 4, 3, 2, 1} with the lowest modelled time per step: the slowest CTA computes r_max + K - 1 rows
    // per step on average (trapezoid) and an exchange costs about 6000 cycles whatever its size (fitted to sweeps over K
    // on C1 / pipe / HPP lattices: 3.2 us per exchange, 2.8 (FHP) / 1.2 (HPP) cycles per word and step).
    ResPlan best;
    memset(&best, 0, sizeof(best));
    double best_cost = 1e300;
    static const int ks[] = {8, 6, 4, 3, 2, 1};
    for (int ki = 0; ki < 6; ++ki) {
        const int K = h->cfg.k_fuse > 0 ? h->cfg.k_fuse : ks[ki];
        if (K > 8) break;
        const int H = hpp ? K : ((K + 1) & ~1);
        const int units = rows / unit;
        int G = std::min(sms, units / std::max(1, (H + unit - 1) / unit)); // every strip at least H rows high
        if (G >= 1) {
            const int base = units / G, extra = units % G;
            const int r_max = (base + (extra ? 1 : 0)) * unit, r_min = base * unit;
            const size_t smem = (size_t)(2 * nd + nm) * (size_t)(r_max + 2 * H) * g.pitch * sizeof(uint32_t);
            if (r_min >= H && smem <= (size_t)RES_MAX_SMEM - 1024) {
                const double cost = (double)(r_max + K - 1) * g.nw * (hpp ? 1.2 : 2.8) + 6000.0 / K;
                if (cost < best_cost) {
                    best_cost = cost;
                    ResPlan& q = best;
                    q.ok = 1; q.G = G; q.unit = unit; q.base_units = base; q.extra_units = extra; q.H = H; q.K = K;
                    q.rows_max = r_max + 2 * H;
                    q.smem_bytes = smem;
                    q.exch_words = (size_t)2 * G * 2 * nd * H * g.pitch * 2; // {word, tag} messages
                    // thread mapping: static ownership while a thread owns at most 2 words of the CTA's local rows (beyond,
                    // the register-resident per-word state spills and the LDS.128 groups win), else dynamic groups
                    const size_t words = (size_t)q.rows_max * g.nw;
                    q.own = words <= 1024 ? 1 : (words <= 2048 ? 2 : 0);
                    if (h->cfg.flags & LGCA_B200_FLAG_RESIDENT_DYNAMIC) q.own = 0;
                }
            }
        }
        if (h->cfg.k_fuse > 0) break;
    }
    if (!best.ok && h->cfg.k_fuse > 0) {
        // the requested depth does not fit: fall back to shallower ones
        for (int K = std::min(h->cfg.k_fuse, 8) - 1; K >= 1 && !best.ok; --K) {
            lgca_b200_lattice tmp = *h;
            tmp.cfg.k_fuse = K;
            best = res_plan(&tmp);
        }
    }
    return best;
}

bool resident_supported(const lgca_b200_lattice* h) { return res_plan(h).ok != 0; }

template <int MODEL, bool NS, bool SL, int OWN>
static int launch_res_own(lgca_b200_lattice* h, const ResPlan& p, ResArgs& A, cudaStream_t s, bool prepare_only)
{
    auto kernel = step_resident_kernel<MODEL, NS, SL, OWN>;
    LGCA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_MAX_SMEM - 1024));
    if (prepare_only) return 0;
    void* args[] = {(void*)&A};
    LGCA_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)kernel, dim3(p.G, 1, 1), dim3(ResThreads<OWN>::value, 1, 1), args,
                                                p.smem_bytes, s));
    h->launches++;
    return 0;
}

template <int MODEL, bool NS, bool SL>
static int launch_res_variant(lgca_b200_lattice* h, const ResPlan& p, ResArgs& A, cudaStream_t s, bool prepare_only)
{
#define RES_GO(OWN) launch_res_own<MODEL, NS, SL, OWN>(h, p, A, s, prepare_only)
    if (prepare_only) { // every mapping gets its shared-memory opt-in
        int rc = RES_GO(0);
        if (!rc) rc = RES_GO(1);
        if (!rc) rc = RES_GO(2);
        return rc;
    }
    switch (p.own) {
    case 1: return RES_GO(1);
    case 2: return RES_GO(2);
    default: return RES_GO(0);
    }
#undef RES_GO
}

template <int MODEL>
static int launch_res_model(lgca_b200_lattice* h, const ResPlan& p, ResArgs& A, cudaStream_t s, bool prep)
{
    if (h->has_sl) return h->has_ns ? launch_res_variant<MODEL, true, true>(h, p, A, s, prep) : launch_res_variant<MODEL, false, true>(h, p, A, s, prep);
    return h->has_ns ? launch_res_variant<MODEL, true, false>(h, p, A, s, prep) : launch_res_variant<MODEL, false, false>(h, p, A, s, prep);
}

// n_steps updates in ONE launch: in -> out.  Returns LGCA_B200_ESTATE when the lattice does not fit on chip.
int launch_step_resident(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int n_steps, cudaStream_t s)
{
    const ResPlan p = res_plan(h);
    if (!p.ok) return set_error(LGCA_B200_ESTATE, "lattice does not fit the SM-resident kernel");
    const Geom& g = h->g;
    if (h->res_exch_words < p.exch_words) {
        // (re)allocation: nothing of this handle is in flight on the buffer after a stream sync
        LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
        if (h->res_exch) { cudaFree(h->res_exch); h->device_bytes -= h->res_exch_words * sizeof(uint32_t); }
        h->res_exch = nullptr; h->res_exch_words = 0;
        LGCA_CUDA_CHECK(cudaMalloc((void**)&h->res_exch, p.exch_words * sizeof(uint32_t)));
        LGCA_CUDA_CHECK(cudaMemset(h->res_exch, 0, p.exch_words * sizeof(uint32_t))); // tag 0 = "nothing here"
        LGCA_CUDA_CHECK(cudaDeviceSynchronize());
        h->res_exch_words = p.exch_words;
        h->res_epoch = 0;
        h->device_bytes += p.exch_words * sizeof(uint32_t);
    }
    ResArgs A;
    A.in = in; A.out = out; A.ns = h->ns; A.sl = h->sl; A.ch = h->ch; A.xedge = h->xedge;
    A.exch = (uint2*)h->res_exch; A.epoch_base = h->res_epoch;
    A.rows = g.rows; A.pitch = g.pitch; A.nw = g.nw; A.rem = g.rem; A.dim_y_south = g.row_south; A.dim_y_north = g.row_north;
    A.plane_stride = (uint32_t)g.plane_stride;
    A.G = p.G; A.unit = p.unit; A.base_units = p.base_units; A.extra_units = p.extra_units; A.H = p.H; A.K = p.K;
    A.rows_max = p.rows_max; A.n_steps = n_steps;
    int rc;
    switch (rule_of(h->cfg.model)) {
    case MODEL_HPP:   rc = launch_res_model<MODEL_HPP>(h, p, A, s, in == nullptr); break;
    case MODEL_FHP_I: rc = launch_res_model<MODEL_FHP_I>(h, p, A, s, in == nullptr); break;
    default:          rc = launch_res_model<MODEL_FHP_II>(h, p, A, s, in == nullptr); break;
    }
    if (rc) return rc;
    if (in) h->res_epoch += (uint32_t)((n_steps + p.K - 1) / p.K); // counters are monotonic across launches
#ifdef LGCA_RES_TIMING
    if (in && n_steps >= 100) {
        static unsigned long long host_t[160][4];
        cudaStreamSynchronize(s);
        cudaMemcpyFromSymbol(host_t, g_res_timing, sizeof(host_t));
        unsigned long long z[160][4] = {};
        cudaMemcpyToSymbol(g_res_timing, z, sizeof(z));
        fprintf(stderr, "res timing G=%d K=%d H=%d own=%d steps=%d: cycles/step of CTA 0 / G/2 / G-1: ", p.G, p.K, p.H, p.own, n_steps);
        for (int j : {0, p.G / 2, p.G - 1})
            fprintf(stderr, "[poll %.0f steps %.0f publish %.0f all %.0f] ", (double)host_t[j][0] / n_steps, (double)host_t[j][1] / n_steps,
                    (double)host_t[j][2] / n_steps, (double)host_t[j][3] / n_steps);
        fprintf(stderr, "\n");
    }
#endif
    return 0;
}

int resident_info(const lgca_b200_lattice* h, int* ctas, int* steps_per_exchange, size_t* smem_bytes)
{
    const ResPlan p = res_plan(h);
    if (ctas) *ctas = p.ok ? p.G : 0;
    if (steps_per_exchange) *steps_per_exchange = p.ok ? p.K : 0;
    if (smem_bytes) *smem_bytes = p.ok ? p.smem_bytes : 0;
    return p.ok;
}

} // namespace lgca_b200
