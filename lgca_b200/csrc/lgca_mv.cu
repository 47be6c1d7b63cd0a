// Order-exact mean velocity: the reference's ONE-THREAD result of OMP_Lattice<M>::get_mean_velocity
// (src/omp_lattice.cpp:508-557) without a sequential pass over the cells.
//
// The reference adds m_x/rho and m_y/rho of every FLUID cell, in cell order, into two float32 accumulators.  On the
// app lattices (1e6 .. 1e7 cells) those sums grow far beyond 2^20, where one float32 ulp is a sizeable fraction of every
// addend: the printed mean velocity -- and with it the forcing decisions of the pipe / Karman schedule
// (apps/karman/karman_viewer.cpp:118-127) -- depends on every single rounding, i.e. on the ORDER.  Reproducing the
// digits therefore needs the sequential semantics, but not a sequential machine:
//
//   * a cell's addend is a function of its 7-bit state byte (128 classes; v = fl(fl(m)/fl(rho)), table built on the
//     host with the reference's own float operations);
//   * while the accumulator s stays inside one binade [2^e, 2^(e+1)) its ulp U = 2^(e-23) is fixed, s = S*U with an
//     integer S, and  fl(s + v) = (S + n)*U  where n = v/U rounded to nearest -- an INTEGER addition; exact ties
//     (v/U = n + 1/2) round to even and so depend on the parity of S: the add is a 2-state automaton on parity(S);
//   * integer additions compose: for a SEGMENT of cells, a binade e and an incoming parity p the device computes
//     { delta, lo, hi } = the sum of the increments and the smallest / largest prefix sum (mv_summary_kernel: one
//     thread per (component, parity, binade), 1024 cells per block, all integer work);
//   * the host walks the segments in order (mv_walk): if S + lo and S + hi stay strictly inside the binade, the whole
//     segment is one integer add; otherwise (a binade boundary or zero is crossed inside the segment: a few segments
//     for the drifting x sum, the random-walk share for y) the segment's cells are added one by one from their class
//     bytes with real float32 additions.
//
// Everything is computed from the SNAPSHOT (the buffer post_process reads), on the post-processing stream.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "lgca_internal.h"

namespace lgca_b200 {

constexpr int     MV_E_MIN     = -3;        // binades [2^-3, 2^25): below, the accumulator is walked cell by cell
constexpr int     MV_NE        = 28;
constexpr int     MV_SEG_WORDS = 32;        // words (of 32 sites) per segment; segments never span rows
constexpr int     MV_SEG_CELLS = MV_SEG_WORDS * 32;
constexpr int32_t MV_SAT       = 1 << 24;   // increments beyond +-2^24 ulps leave the binade anyway
constexpr int     MV_SLOTS     = 4 * MV_NE; // (component, parity, binade) trajectories per segment
constexpr int     MV_REC       = MV_SLOTS * 3; // int32 per segment
// Records are stored slot-major: { delta, lo, hi } of slot (component*2 + parity)*MV_NE + binade for segment `seg` of
// `nseg`.  The walk stays in one binade for long runs of segments, so it reads two (parity) sequential streams.
__host__ __device__ __forceinline__ size_t mv_rec_index(int slot, size_t seg, size_t nseg) { return ((size_t)slot * nseg + seg) * 3; }

struct MvTables {
    int32_t r[2][128][MV_NE];  // (n << 1) | tie  for component, state byte, binade (device layout: lanes = binades)
    int32_t rt[2][MV_NE][128]; // the same, binade-major (host walk)
    int32_t has_tie[2][MV_NE]; // any class with an exact tie in this binade?
    float   v[2][128];         // the addend itself
};

// float(sin(M_PI/3)): `static constexpr Real SIN`, src/lgca_models.h:236
static const float MV_SIN = 0.866025388f;

static void build_tables(int model, MvTables& t)
{
    const int   nd = model == LGCA_B200_HPP ? 4 : (model == LGCA_B200_FHP_I ? 6 : 7);
    const float vx4[4] = {1.0f, 0.0f, -1.0f, 0.0f}, vy4[4] = {0.0f, 1.0f, 0.0f, -1.0f};
    const float vx7[7] = {1.0f, 0.5f, -0.5f, -1.0f, -0.5f, 0.5f, 0.0f};
    const float vy7[7] = {0.0f, MV_SIN, MV_SIN, 0.0f, -MV_SIN, -MV_SIN, 0.0f};
    memset(&t, 0, sizeof(t));
    for (int b = 0; b < (1 << nd); ++b) {
        // cell_post_process, src/omp_lattice.cpp:360-394: char density, float momenta accumulated direction by direction
        volatile float mx = 0.0f, my = 0.0f;
        int dens = 0;
        for (int d = 0; d < nd; ++d) {
            const int ns = (b >> d) & 1;
            dens += ns;
            mx = mx + (float)ns * (nd == 4 ? vx4[d] : vx7[d]);
            my = my + (float)ns * (nd == 4 ? vy4[d] : vy7[d]);
        }
        if (dens == 0) continue; // `cell_density > 1.0e-06` fails: nothing is added
        volatile float v[2];
        v[0] = mx / (float)dens;
        v[1] = my / (float)dens;
        for (int c = 0; c < 2; ++c) {
            t.v[c][b] = v[c];
            for (int e = 0; e < MV_NE; ++e) {
                // q = v / U is exact in double (v has 24 significant bits, U is a power of two)
                const double q = ldexp((double)v[c], 23 - (MV_E_MIN + e));
                double fl = floor(q);
                const double frac = q - fl;
                int tie = 0;
                if (frac == 0.5) tie = 1;
                else if (frac > 0.5) fl += 1.0;
                if (fl > (double)MV_SAT) { fl = (double)MV_SAT; tie = 0; }
                if (fl < -(double)MV_SAT) { fl = -(double)MV_SAT; tie = 0; }
                t.r[c][b][e] = t.rt[c][e][b] = (int32_t)fl * 2 + tie;
                t.has_tie[c][e] |= tie;
            }
        }
    }
}

const MvTables& mv_tables(int model)
{
    static MvTables       tabs[4];
    static std::once_flag once[4];
    std::call_once(once[model], [model] { build_tables(model, tabs[model]); });
    return tabs[model];
}

// one trajectory step: S (as offset `acc` from the incoming value of parity `par`) takes the addend coded in r
__host__ __device__ __forceinline__ void mv_add(int32_t r, int par, int32_t& acc, int32_t& lo, int32_t& hi)
{
    const int32_t n = r >> 1;
    // tie: S + n + 1/2 rounds to the even neighbour
    acc += n + (r & (par ^ acc ^ n) & 1);
    acc = acc > (1 << 28) ? (1 << 28) : (acc < -(1 << 28) ? -(1 << 28) : acc);
    lo = acc < lo ? acc : lo;
    hi = acc > hi ? acc : hi;
}

// ---- device: class bytes + per-segment summaries ------------------------------------------------------------------
// grid = segments of the owned rows; block = 128 threads: warp = (component, parity), lane = binade.
template <int ND>
__global__ void __launch_bounds__(128) mv_summary_kernel(const uint32_t* __restrict__ planes, const uint32_t* __restrict__ ns,
                                                         const uint32_t* __restrict__ sl, const int32_t* __restrict__ tab,
                                                         int32_t* __restrict__ rec, uint32_t* __restrict__ fluid,
                                                         uint8_t* __restrict__ cls, const Geom g, uint32_t segs_per_row, uint32_t nseg)
{
    __shared__ int32_t s_tab[2 * 128 * MV_NE];
    __shared__ uint8_t s_cls[MV_SEG_CELLS];
    __shared__ uint32_t s_fluid;
    const uint32_t seg = blockIdx.x, row = seg / segs_per_row, k = seg % segs_per_row;
    const uint32_t x0 = k * MV_SEG_CELLS;
    const uint32_t ncells = min((uint32_t)MV_SEG_CELLS, g.dim_x - x0);
    const int t = threadIdx.x;
    if (t == 0) s_fluid = 0;
    for (int i = t; i < 2 * (1 << ND) * MV_NE; i += 128) { // only the model's 2^ND classes are ever looked up
        const int c = i / ((1 << ND) * MV_NE), rest = i % ((1 << ND) * MV_NE);
        s_tab[c * 128 * MV_NE + rest] = __ldg(tab + c * 128 * MV_NE + rest);
    }
    __syncthreads();
    {   // phase 1: state byte of 8 sites per thread (0 for solid cells and beyond the row end)
        const uint32_t w = k * MV_SEG_WORDS + (t >> 2);
        const int      b0 = (t & 3) * 8;
        uint32_t v[ND], solid = 0xFFFFFFFFu;
        if (w < g.nw) {
            const size_t base = (size_t)(row + g.halo) * g.pitch + w;
#pragma unroll
            for (int d = 0; d < ND; ++d) v[d] = __ldg(planes + (size_t)d * g.plane_stride + base);
            solid = __ldg(ns + base) | __ldg(sl + base) | ~valid_mask(g, (int)w);
        } else {
#pragma unroll
            for (int d = 0; d < ND; ++d) v[d] = 0;
        }
        uint32_t nf = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int bit = b0 + j;
            uint32_t b = 0;
#pragma unroll
            for (int d = 0; d < ND; ++d) b |= ((v[d] >> bit) & 1u) << d;
            const uint32_t f = ((solid >> bit) & 1u) ^ 1u;
            nf += f;
            s_cls[(t >> 2) * 32 + bit] = (uint8_t)(f ? b : 0u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nf += __shfl_down_sync(0xFFFFFFFFu, nf, o);
        if ((t & 31) == 0 && nf) atomicAdd(&s_fluid, nf);
    }
    __syncthreads();
    const size_t cell0 = (size_t)row * g.dim_x + x0;
    for (uint32_t i = t; i < ncells; i += 128) cls[cell0 + i] = s_cls[i];
    if (t == 0) fluid[seg] = s_fluid;
    // phase 2: one trajectory per thread
    const int lane = t & 31, warp = t >> 5;
    if (lane < MV_NE) {
        const int comp = warp >> 1, par = warp & 1;
        const int32_t* tb = s_tab + comp * 128 * MV_NE + lane;
        int32_t acc = 0, lo = 0x7FFFFFFF, hi = (int32_t)0x80000000;
        for (uint32_t i = 0; i < ncells; ++i) {
            const uint32_t c = s_cls[i];
            if (c == 0) continue; // uniform: every lane reads the same class
            mv_add(tb[c * MV_NE], par, acc, lo, hi);
        }
        int32_t* out = rec + mv_rec_index(warp * MV_NE + lane, seg, nseg);
        out[0] = acc; out[1] = lo; out[2] = hi;
    }
}

// ---- host: the same summaries (CPU tests of the algorithm; lgca_b200_mean_velocity_replay) ---------------------------
static void mv_summarise_host(const MvTables& T, const uint8_t* cls, size_t ncells, int32_t* rec, size_t seg, size_t nseg)
{
    for (int w = 0; w < 4; ++w)
        for (int e = 0; e < MV_NE; ++e) {
            int32_t acc = 0, lo = 0x7FFFFFFF, hi = (int32_t)0x80000000;
            for (size_t i = 0; i < ncells; ++i)
                if (cls[i]) mv_add(T.r[w >> 1][cls[i]][e], w & 1, acc, lo, hi);
            int32_t* out = rec + mv_rec_index(w * MV_NE + e, seg, nseg);
            out[0] = acc; out[1] = lo; out[2] = hi;
        }
}

// ---- host: the ordered walk over one strip's segments ----------------------------------------------------------------
// sums[c] continues from its incoming value.  stats (optional): [0] segments taken as one integer add, [1] walked.
static void mv_walk_component(const MvTables& T, const int32_t* rec, const uint8_t* cls, uint32_t dim_x, uint32_t rows, int c,
                              float* sum, uint64_t* stats)
{
    const uint32_t spr = (dim_x + MV_SEG_CELLS - 1) / MV_SEG_CELLS;
    const size_t   nseg = (size_t)spr * rows;
    const int64_t  LO = (int64_t)1 << 23, HI = (int64_t)1 << 24;
    // The accumulator lives in one of two forms: a float32 `a` (below 2^MV_E_MIN, at zero, beyond the tabulated binades),
    // or -- for as long as it stays inside one binade -- the pair (e, S) with a = S * 2^(e-23), |S| in [2^23, 2^24).  In
    // the second form a whole segment is one integer add and nothing is converted in between (the float32 value is
    // formed again only where the accumulator leaves its binade: plain IEEE float32 additions, x86-64 SSE scalar adds;
    // the host code is built without -ffast-math).
    float    a = *sum;
    int      e = 0, idx = -1; // idx >= 0: integer form
    int64_t  S = 0;
    auto to_integer = [&]() {
        idx = -1;
        if (a != 0.0f && isfinite(a)) {
            e = ilogbf(fabsf(a));
            const int i = e - MV_E_MIN;
            if (i >= 0 && i < MV_NE) { idx = i; S = (int64_t)ldexpf(a, 23 - e); } // exact
        }
    };
    auto to_float = [&]() { a = ldexpf((float)S, e - 23); idx = -1; };
    uint64_t n_fast = 0, n_walked = 0;
    to_integer();
    for (uint32_t row = 0; row < rows; ++row)
        for (uint32_t k = 0; k < spr; ++k) {
            const size_t seg = (size_t)row * spr + k;
            if (idx < 0) to_integer();
            if (idx >= 0) {
                const int32_t* q = rec + mv_rec_index((c * 2 + (int)(S & 1)) * MV_NE + idx, seg, nseg);
                const int64_t lo = q[1], hi = q[2];
                if (hi < lo) { ++n_fast; continue; } // no addend in this segment
                if (S > 0 ? (S + lo > LO && S + hi < HI) : (S + hi < -LO && S + lo > -HI)) {
                    S += q[0];
                    ++n_fast;
                    continue;
                }
            }
            ++n_walked;
            // a binade boundary (or zero) is crossed inside this segment: the reference's own float32 additions, cell by
            // cell.  (Measured on the host: integer adds inside the binade with a conversion at every exit were SLOWER here
            // than the plain 4-cycle add chain -- the walked segments are mostly those of the undriven component close to
            // zero, where the binades are narrow and almost every add leaves them: 3.4 -> 1.5 ms on the Karman default.)
            if (idx >= 0) to_float();
            const uint32_t x0 = k * MV_SEG_CELLS, n = std::min<uint32_t>(MV_SEG_CELLS, dim_x - x0);
            const uint8_t* p = cls + (size_t)row * dim_x + x0;
            const float*   tv = T.v[c];
            for (uint32_t i = 0; i < n; ++i) a += tv[p[i]];
        }
    if (idx >= 0) to_float();
    *sum = a;
    if (stats) { stats[0] += n_fast; stats[1] += n_walked; }
}

// the two components are independent chains: y walks on a helper thread while x walks here
static void mv_walk(const MvTables& T, const int32_t* rec, const uint8_t* cls, uint32_t dim_x, uint32_t rows, float sums[2],
                    uint64_t* stats)
{
    uint64_t st[2][2] = {{0, 0}, {0, 0}};
    if ((size_t)dim_x * rows >= 200000) {
        std::thread ty([&] { mv_walk_component(T, rec, cls, dim_x, rows, 1, sums + 1, st[1]); });
        mv_walk_component(T, rec, cls, dim_x, rows, 0, sums + 0, st[0]);
        ty.join();
    } else {
        for (int c = 0; c < 2; ++c) mv_walk_component(T, rec, cls, dim_x, rows, c, sums + c, st[c]);
    }
    if (stats) { stats[0] += st[0][0] + st[1][0]; stats[1] += st[0][1] + st[1][1]; }
}

static int ensure_mv_buffers(lgca_b200_lattice* h, size_t nseg, size_t cells)
{
    if (h->d_mv_rec) return 0;
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_mv_tab, sizeof(int32_t) * 2 * 128 * MV_NE));
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_mv_rec, nseg * MV_REC * sizeof(int32_t)));
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_mv_fluid, nseg * sizeof(uint32_t)));
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_mv_cls, cells));
    h->device_bytes += sizeof(int32_t) * 2 * 128 * MV_NE + nseg * MV_REC * sizeof(int32_t) + nseg * sizeof(uint32_t) + cells;
    LGCA_CUDA_CHECK(cudaHostAlloc((void**)&h->h_mv_rec, nseg * MV_REC * sizeof(int32_t), cudaHostAllocDefault));
    LGCA_CUDA_CHECK(cudaHostAlloc((void**)&h->h_mv_fluid, nseg * sizeof(uint32_t), cudaHostAllocDefault));
    LGCA_CUDA_CHECK(cudaHostAlloc((void**)&h->h_mv_cls, cells, cudaHostAllocDefault));
    const MvTables& T = mv_tables(h->cfg.model);
    LGCA_CUDA_CHECK(cudaMemcpy(h->d_mv_tab, T.r, sizeof(T.r), cudaMemcpyHostToDevice));
    return 0;
}

void free_mv_buffers(lgca_b200_lattice* h)
{
    cudaFree(h->d_mv_tab); cudaFree(h->d_mv_rec); cudaFree(h->d_mv_fluid); cudaFree(h->d_mv_cls);
    if (h->h_mv_rec) cudaFreeHost(h->h_mv_rec);
    if (h->h_mv_fluid) cudaFreeHost(h->h_mv_fluid);
    if (h->h_mv_cls) cudaFreeHost(h->h_mv_cls);
}

} // namespace lgca_b200

using namespace lgca_b200;

extern "C" {

int lgca_b200_mean_velocity_exact(lgca_b200_lattice* h, float sums[2], uint64_t* fluid_cells)
{
    if (!h || !sums || !fluid_cells) return set_error(LGCA_B200_EINVAL, "null argument");
    const Geom& g = h->g;
    const uint32_t own = g.rows - 2 * g.halo;
    const size_t   cells = (size_t)g.dim_x * own;
    if (cells > ((size_t)1 << 28))
        return set_error(LGCA_B200_EINVAL, "order-exact mean velocity is limited to 2^28 cells per handle (float32 sums saturate "
                                           "long before); use lgca_b200_mean_velocity");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    const uint32_t spr = (g.dim_x + MV_SEG_CELLS - 1) / MV_SEG_CELLS;
    const size_t   nseg = (size_t)spr * own;
    int rc = ensure_mv_buffers(h, nseg, cells);
    if (rc) return rc;
    cudaStream_t s = h->s_post;
    auto now_ns = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec; };
    const uint64_t t0 = now_ns();
    {
        SnapLock lock(h);
        LGCA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_snap, 0));
#define MV(ND) mv_summary_kernel<ND><<<(unsigned)nseg, 128, 0, s>>>(h->snap, h->ns, h->sl, h->d_mv_tab, h->d_mv_rec, h->d_mv_fluid, \
                                                                      h->d_mv_cls, g, spr, (uint32_t)nseg)
        if (h->nd == 4) MV(4); else if (h->nd == 6) MV(6); else MV(7);
#undef MV
        h->launches++;
        LGCA_CUDA_CHECK(cudaGetLastError());
        LGCA_CUDA_CHECK(cudaMemcpyAsync(h->h_mv_rec, h->d_mv_rec, nseg * MV_REC * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        LGCA_CUDA_CHECK(cudaMemcpyAsync(h->h_mv_fluid, h->d_mv_fluid, nseg * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        LGCA_CUDA_CHECK(cudaMemcpyAsync(h->h_mv_cls, h->d_mv_cls, cells, cudaMemcpyDeviceToHost, s));
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_post, s));
    }
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    const uint64_t t1 = now_ns();
    uint64_t nf = 0;
    for (size_t i = 0; i < nseg; ++i) nf += h->h_mv_fluid[i];
    *fluid_cells += nf;
    mv_walk(mv_tables(h->cfg.model), h->h_mv_rec, h->h_mv_cls, g.dim_x, own, sums, h->mv_stats);
    h->mv_stats[2] += t1 - t0;
    h->mv_stats[3] += now_ns() - t1;
    return 0;
}

int lgca_b200_mean_velocity_stats(lgca_b200_lattice* h, uint64_t out[4])
{
    if (!h || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    for (int i = 0; i < 4; ++i) out[i] = h->mv_stats[i];
    return 0;
}

int lgca_b200_mean_velocity_replay(int model, const uint8_t* class_bytes, uint32_t dim_x, uint32_t rows, float sums[2],
                                   uint64_t* segments_fast, uint64_t* segments_walked)
{
    if (model < LGCA_B200_HPP || model > LGCA_B200_FHP_III || (!class_bytes && dim_x && rows) || !sums)
        return set_error(LGCA_B200_EINVAL, "bad argument");
    const MvTables& T = mv_tables(model);
    const uint32_t spr = (dim_x + MV_SEG_CELLS - 1) / MV_SEG_CELLS;
    const size_t nseg = (size_t)spr * rows;
    std::vector<int32_t> rec(nseg * MV_REC);
    for (uint32_t row = 0; row < rows; ++row)
        for (uint32_t k = 0; k < spr; ++k) {
            const uint32_t x0 = k * MV_SEG_CELLS, n = std::min<uint32_t>(MV_SEG_CELLS, dim_x - x0);
            mv_summarise_host(T, class_bytes + (size_t)row * dim_x + x0, n, rec.data(), (size_t)row * spr + k, nseg);
        }
    uint64_t stats[2] = {0, 0};
    mv_walk(T, rec.data(), class_bytes, dim_x, rows, sums, stats);
    if (segments_fast) *segments_fast = stats[0];
    if (segments_walked) *segments_walked = stats[1];
    return 0;
}

} // extern "C"
