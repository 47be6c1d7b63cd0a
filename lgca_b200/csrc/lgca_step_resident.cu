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
//   * Every K steps the CTAs exchange ghost rows through L2: each CTA bulk-stores its top and bottom H rows into an
//     exchange area (cp.async.bulk shared->global), publishes a counter with release semantics, and its two neighbours
//     acquire it and bulk-load the rows into their ghost rows.  Only NEIGHBOURS synchronise -- no grid-wide barrier.
//     The exchange area is double-buffered by block parity; the counter protocol orders every reuse (same argument as
//     the multi-GPU ring, lgca_ring.cu).
//   * Row ends: words are row-aligned here, so widths that are not a multiple of 32 only change where the carry bit
//     of the x-shift comes from at the first / last word of a row.
#include <string.h>

#include <algorithm>

#include "lgca_internal.h"

namespace lgca_b200 {

constexpr int RES_THREADS = 1024;
constexpr int RES_MAX_SMEM = 227 * 1024; // opt-in dynamic shared memory per CTA on sm_100

struct ResArgs {
    const uint32_t* in;       // [nd][rows][pitch]
    uint32_t*       out;
    const uint32_t* ns;
    const uint32_t* sl;
    const uint32_t* ch;
    const uint32_t* xedge;
    uint32_t*       exch;     // [2 parities][G][2 sides][nd][H][pitch]
    uint32_t*       flags;    // [G] blocks published by CTA j (monotonic over launches: + epoch_base)
    uint32_t        epoch_base;
    uint32_t        rows, pitch, nw, rem, dim_y_south, dim_y_north; // rows of the lattice; stored rows of the N/S domain edges
    uint32_t        plane_stride;  // words between planes in global memory
    int             G;        // CTAs
    int             unit;     // strip heights are multiples of unit (2 for the hexagonal models)
    int             base_units, extra_units; // CTA j owns (base + (j < extra)) units
    int             H, K;     // ghost rows per side, steps per exchange
    int             rows_max; // rows of the shared-memory buffers = max strip height + 2H
    int             n_steps;
};

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

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// first row and height of CTA j's strip
__device__ __host__ __forceinline__ void res_strip(const ResArgs& A, int j, int& y0, int& R)
{
    const int before = j * A.base_units + (j < A.extra_units ? j : A.extra_units);
    y0 = before * A.unit;
    R  = (A.base_units + (j < A.extra_units ? 1 : 0)) * A.unit;
}

template <int MODEL, bool HAS_NS, bool HAS_SL>
__global__ void __launch_bounds__(RES_THREADS, 1) step_resident_kernel(const ResArgs A)
{
    constexpr int  ND  = num_dir_of(MODEL);
    constexpr bool HPP = rule_of(MODEL) == MODEL_HPP;
    constexpr int  NM  = (HPP ? 0 : 1) + (HAS_NS ? 1 : 0) + (HAS_SL ? 1 : 0); // static mask planes held on chip
    extern __shared__ __align__(128) uint32_t smem[];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
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

    // ---- stage the strip + ghost rows (periodic in y) and the masks into shared memory --------------------------
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
    // rows are copied with their padding words (pitch > nw): keep them zero in the second buffer too (the first one got
    // zeros from global memory), they travel to the exchange area and back to global memory with the rows
    if (P > nw) {
        const int pad = P - nw;
        for (int e = tid; e < A.rows_max * pad * ND; e += RES_THREADS) {
            const int d = e / (A.rows_max * pad), q = e % (A.rows_max * pad);
            buf1[d * plane_sz + (q / pad) * P + nw + (q % pad)] = 0u;
        }
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    __syncthreads();

    // thread -> (row, word) stepping: element index e = tid, tid + T, ...; (dq, dr) = divmod(T, nw)
    const int dq = RES_THREADS / nw, dr = RES_THREADS % nw;
    const int q0 = tid / nw, w0 = tid % nw;

    uint32_t* cur = buf0;
    uint32_t* nxt = buf1;
    const int n_blocks = (A.n_steps + A.K - 1) / A.K;
    const int lower = (j + A.G - 1) % A.G, upper = (j + 1) % A.G;
    const size_t side_words = (size_t)ND * H * P;                  // one side of one CTA in the exchange area
    const size_t parity_words = (size_t)A.G * 2 * side_words;

    for (int b = 0; b < n_blocks; ++b) {
        const int kb = min(A.K, A.n_steps - b * A.K);
        if (b > 0) {
            // ghost rows of this block: the neighbours' edge rows after block b-1
            if (tid == 0) {
                const uint32_t want = A.epoch_base + (uint32_t)b;
                while ((int32_t)(ld_acquire_gpu(A.flags + lower) - want) < 0) { }
                while ((int32_t)(ld_acquire_gpu(A.flags + upper) - want) < 0) { }
                fence_proxy_async();
                const uint32_t* ex = A.exch + (size_t)((b - 1) & 1) * parity_words;
                const uint32_t bytes = (uint32_t)H * P * 4u;
                mbar_expect_tx(&bar, 2u * ND * bytes);
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    // my lower ghost rows [0, H) <- the lower neighbour's TOP side; upper ghost rows <- upper neighbour's BOTTOM side
                    bulk_g2s(cur + d * plane_sz, ex + ((size_t)lower * 2 + 1) * side_words + (size_t)d * H * P, bytes, &bar);
                    bulk_g2s(cur + d * plane_sz + (size_t)(H + R) * P, ex + ((size_t)upper * 2 + 0) * side_words + (size_t)d * H * P, bytes, &bar);
                }
            }
            mbar_wait(&bar, phase);
            phase ^= 1;
        }
        for (int s = 1; s <= kb; ++s) {
            // rows that are still needed and still valid after step s of this block
            const int ra = H - (kb - s), rb = H + R + (kb - s);
            const int total = (rb - ra) * nw;
            int r = ra + q0, w = w0;
            for (int e = tid; e < total; e += RES_THREADS) {
                const uint32_t rc = (uint32_t)r * P, rs = rc - P, rn = rc + P;
                const bool first = (w == 0), last = (w == nw - 1);
                const int  wl = first ? nw - 1 : w - 1, wr = last ? 0 : w + 1;
                // carry bit of the 1-bit x-shifts at the row ends (periodic; the last word may be partial)
                const int  lsh = (first && rem) ? 32 - rem : 0;      // pre-shift of the left neighbour word
                const int  hi  = (last && rem) ? rem - 1 : 31;       // bit that receives site 0 in a down-shift
#define S(d, ro, ww) cur[(d) * plane_sz + (ro) + (ww)]
#define UP(d, ro)   __funnelshift_l(S(d, ro, wl) << lsh, S(d, ro, w), 1)
#define DOWN(d, ro) ((S(d, ro, w) >> 1) | ((S(d, ro, wr) & 1u) << hi))
                uint32_t n[7];
                if (HPP) {
                    n[0] = UP(0, rc);
                    n[2] = DOWN(2, rc);
                    n[1] = S(1, rs, w);
                    n[3] = S(3, rn, w);
                    n[4] = n[5] = n[6] = 0u;
                } else {
                    n[0] = UP(0, rc);
                    n[3] = DOWN(3, rc);
                    if (!(r & 1)) { // local parity == global parity (strip starts and H are even)
                        n[1] = UP(1, rs);
                        n[2] = S(2, rs, w);
                        n[4] = S(4, rn, w);
                        n[5] = UP(5, rn);
                    } else {
                        n[1] = S(1, rs, w);
                        n[2] = DOWN(2, rs);
                        n[4] = DOWN(4, rn);
                        n[5] = S(5, rn, w);
                    }
                    n[6] = ND == 7 ? S(6, rc, w) : 0u;
                }
#undef S
#undef UP
#undef DOWN
                const uint32_t p  = HPP ? 0u : m_ch[rc + w];
                const uint32_t ns = HAS_NS ? m_ns[rc + w] : 0u;
                const uint32_t sl = HAS_SL ? m_sl[rc + w] : 0u;
                uint32_t ew = 0u, ns_row = 0u;
                if (HAS_SL) {
                    ew = __ldg(A.xedge + w);
                    int gy = y0 - H + r;                              // global (stored) row of this local row
                    if (gy < 0) gy += (int)A.rows; else if (gy >= (int)A.rows) gy -= (int)A.rows;
                    ns_row = ((uint32_t)gy == A.dim_y_south || (uint32_t)gy == A.dim_y_north) ? 0xFFFFFFFFu : 0u;
                }
                collide_and_walls<MODEL, HAS_NS, HAS_SL>(n, p, ns, sl, ew, ns_row);
                const uint32_t vm = (last && rem) ? ((1u << rem) - 1u) : 0xFFFFFFFFu;
#pragma unroll
                for (int d = 0; d < ND; ++d) nxt[d * plane_sz + rc + w] = n[d] & vm;
                r += dq; w += dr;
                if (w >= nw) { w -= nw; ++r; }
            }
            __syncthreads();
            uint32_t* t = cur; cur = nxt; nxt = t;
        }
        if (b + 1 < n_blocks) {
            // publish my edge rows for the neighbours' next block
            if (tid == 0) {
                fence_proxy_async(); // the st.shared of the last step -> visible to the bulk stores
                uint32_t* ex = A.exch + (size_t)(b & 1) * parity_words + (size_t)j * 2 * side_words;
                const uint32_t bytes = (uint32_t)H * P * 4u;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    bulk_s2g(ex + (size_t)d * H * P, cur + d * plane_sz + (size_t)H * P, bytes);                 // BOTTOM side: rows [H, 2H)
                    bulk_s2g(ex + side_words + (size_t)d * H * P, cur + d * plane_sz + (size_t)R * P, bytes);    // TOP side: rows [R, R+H)
                }
                bulk_commit();
                bulk_wait_all();
                fence_proxy_async();
                st_release_gpu(A.flags + j, A.epoch_base + (uint32_t)b + 1u);
            }
            // (no block-wide sync needed here: the next block starts with the mbarrier wait led by thread 0)
        }
    }
    // ---- write the strip back ------------------------------------------------------------------------------------
    if (tid == 0) {
        fence_proxy_async();
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
    int    G, unit, base_units, extra_units, H, K, rows_max;
    size_t smem_bytes, exch_words;
};

static ResPlan res_plan(const lgca_b200_lattice* h)
{
    ResPlan p;
    memset(&p, 0, sizeof(p));
    const Geom& g = h->g;
    if (g.halo != 0 || !g.wrap_y) return p;                       // whole lattices only
    if (h->cfg.flags & (LGCA_B200_FLAG_SIMPLE_KERNEL | LGCA_B200_FLAG_NO_RESIDENT)) return p;
    const bool hpp = rule_of(h->cfg.model) == MODEL_HPP;
    const int  nd = h->nd, nm = (hpp ? 0 : 1) + (h->has_ns ? 1 : 0) + (h->has_sl ? 1 : 0);
    const int  unit = hpp ? 1 : 2;
    const int  rows = (int)g.rows, sms = h->sm_count > 0 ? h->sm_count : 148;
    if (rows % unit) return p;
    // steps per ghost-row exchange: cfg.k_fuse when given, else the deepest interval whose buffers fit (fewer
    // exchanges, at the price of more redundant ghost-row work)
    for (int K = h->cfg.k_fuse > 0 ? h->cfg.k_fuse : 8; K >= 1; --K) {
        const int H = hpp ? K : ((K + 1) & ~1);
        const int units = rows / unit;
        int G = std::min(sms, units / std::max(1, (H + unit - 1) / unit)); // every strip at least H rows high
        if (G < 1) continue;
        const int base = units / G, extra = units % G;
        const int r_max = (base + (extra ? 1 : 0)) * unit, r_min = base * unit;
        if (r_min < H) continue;
        const size_t smem = (size_t)(2 * nd + nm) * (size_t)(r_max + 2 * H) * g.pitch * sizeof(uint32_t);
        if (smem > (size_t)RES_MAX_SMEM - 1024) continue;
        p.ok = 1; p.G = G; p.unit = unit; p.base_units = base; p.extra_units = extra; p.H = H; p.K = K;
        p.rows_max = r_max + 2 * H;
        p.smem_bytes = smem;
        p.exch_words = (size_t)2 * G * 2 * nd * H * g.pitch;
        return p;
    }
    return p;
}

bool resident_supported(const lgca_b200_lattice* h) { return res_plan(h).ok != 0; }

template <int MODEL, bool NS, bool SL>
static int launch_res_variant(lgca_b200_lattice* h, const ResPlan& p, ResArgs& A, cudaStream_t s, bool prepare_only)
{
    auto kernel = step_resident_kernel<MODEL, NS, SL>;
    LGCA_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RES_MAX_SMEM - 1024));
    if (prepare_only) return 0;
    void* args[] = {(void*)&A};
    LGCA_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)kernel, dim3(p.G, 1, 1), dim3(RES_THREADS, 1, 1), args, p.smem_bytes, s));
    h->launches++;
    return 0;
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
    if (h->res_exch_words < p.exch_words || h->res_flags_n < p.G) {
        // (re)allocation: nothing of this handle is in flight on the buffers after a stream sync
        LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
        cudaFree(h->res_exch); cudaFree(h->res_flags);
        h->res_exch = nullptr; h->res_flags = nullptr; h->res_exch_words = 0; h->res_flags_n = 0;
        LGCA_CUDA_CHECK(cudaMalloc((void**)&h->res_exch, p.exch_words * sizeof(uint32_t)));
        LGCA_CUDA_CHECK(cudaMalloc((void**)&h->res_flags, (size_t)p.G * sizeof(uint32_t)));
        LGCA_CUDA_CHECK(cudaMemset(h->res_flags, 0, (size_t)p.G * sizeof(uint32_t)));
        LGCA_CUDA_CHECK(cudaDeviceSynchronize());
        h->res_exch_words = p.exch_words; h->res_flags_n = p.G;
        h->res_epoch = 0;
        h->device_bytes += p.exch_words * sizeof(uint32_t);
    }
    ResArgs A;
    A.in = in; A.out = out; A.ns = h->ns; A.sl = h->sl; A.ch = h->ch; A.xedge = h->xedge;
    A.exch = h->res_exch; A.flags = h->res_flags; A.epoch_base = h->res_epoch;
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
