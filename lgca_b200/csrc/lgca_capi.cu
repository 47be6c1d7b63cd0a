// C-ABI of liblgca_b200.so (include/lgca_b200.h): handle management, host<->device plumbing and the
// orchestration of the kernels in the sibling .cu files.  No CPU fallback: every compute entry point
// needs a CUDA device of compute capability 10.x (the library is built for sm_100a only).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "lgca_internal.h"

namespace lgca_b200 {

static thread_local std::string g_last_error;

int set_error(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int set_cuda_error(cudaError_t e, const char* what, const char* file, int line)
{
    return set_error(e == cudaErrorMemoryAllocation ? LGCA_B200_ENOMEM : LGCA_B200_ECUDA, "CUDA error %d (%s) at %s:%d: %s",
                     (int)e, cudaGetErrorString(e), file, line, what);
}

static const size_t STAGE_BYTES = 64u << 20;

static int dev_alloc(lgca_b200_lattice* h, void** p, size_t bytes, bool zero)
{
    LGCA_CUDA_CHECK(cudaMalloc(p, bytes));
    h->device_bytes += bytes;
    if (zero) LGCA_CUDA_CHECK(cudaMemset(*p, 0, bytes));
    return 0;
}

static int ensure_stage(lgca_b200_lattice* h)
{
    if (h->d_stage[0]) return 0;
    const size_t cells = (size_t)h->g.dim_x * (h->g.rows - 2 * h->g.halo);
    size_t want = std::min(STAGE_BYTES, std::max<size_t>(cells * 4, 4096));
    want = std::max(want, (size_t)h->g.dim_x * 4 + 64); // at least one row of int32 cell types
    for (int i = 0; i < 2; ++i) {
        int rc = dev_alloc(h, &h->d_stage[i], want, false);
        if (rc) return rc;
    }
    h->stage_bytes = want;
    return 0;
}

static int ensure_copy_stream(lgca_b200_lattice* h)
{
    if (h->s_copy) return 0;
    LGCA_CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        LGCA_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_stage_free[i], cudaEventDisableTiming));
        LGCA_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_stage_full[i], cudaEventDisableTiming));
    }
    return 0;
}

static inline uint32_t own_rows(const lgca_b200_lattice* h) { return h->g.rows - 2 * h->g.halo; }
static inline size_t plane_words(const lgca_b200_lattice* h) { return (size_t)h->g.plane_stride; }

} // namespace lgca_b200

namespace lgca_b200 {
// Copy-on-write for the zero-copy snapshot: before the live buffer is modified in place (upload, init, body force,
// halo import) the live state moves into the spare buffer and the snapshot keeps the old one.
int unalias_snapshot(lgca_b200_lattice* h)
{
    if (!h->snap_spare) return 0;
    SnapLock lock(h);
    // the spare was the previous snapshot: the post stream may still be reading it
    LGCA_CUDA_CHECK(cudaStreamWaitEvent(h->s_compute, h->ev_post, 0));
    // strips: the copy takes the ghost rows along, so the neighbours' pushes of the current epoch must have landed
    if (h->ring_connected && h->ring_epoch) {
        int rc = ring_wait_current_epoch(h);
        if (rc) return rc;
    }
    LGCA_CUDA_CHECK(cudaMemcpyAsync(h->snap_spare, h->planes[h->cur], (size_t)h->g.plane_stride * sizeof(uint32_t) * h->nd,
                                    cudaMemcpyDeviceToDevice, h->s_compute));
    h->planes[h->cur] = h->snap_spare;
    h->snap_spare = nullptr;
    return 0;
}

// Steps that ONE kernel launch can advance this handle (<= want).  A strip's ghost rows are refreshed only between
// launches, so this is also the number of steps a strip may take per halo exchange: the fused depth the wavefront
// kernel supports for this geometry, or 1 when only the generic kernel applies (dim_x < 64, fewer than 8 stored rows,
// LGCA_B200_FLAG_SIMPLE_KERNEL).
int steps_per_launch(const lgca_b200_lattice* h, int want)
{
    if (want < 1) return 1;
    if (h->cfg.flags & LGCA_B200_FLAG_SIMPLE_KERNEL) return 1;
    int k = std::min(h->k_fuse, want);
    while (k > 1 && !wave_supported(h, k)) --k;
    return std::max(k, 1);
}
} // namespace lgca_b200

namespace lgca_b200 {
// { sum of m_x/rho, sum of m_y/rho, number of FLUID cells } over the owned rows of the snapshot
int mean_velocity_sums(lgca_b200_lattice* h, double out3[3])
{
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->s_post;
    SnapLock lock(h);
    LGCA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_snap, 0));
    int rc = launch_mean_velocity(h, h->snap, h->d_scalars, s);
    if (rc) return rc;
    LGCA_CUDA_CHECK(cudaMemcpyAsync(h->h_scalars, h->d_scalars, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_post, s));
    lock.release();
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    for (int i = 0; i < 3; ++i) out3[i] = h->h_scalars[i];
    return 0;
}
} // namespace lgca_b200

using namespace lgca_b200;

extern "C" {

const char* lgca_b200_last_error(void) { return g_last_error.c_str(); }
int lgca_b200_version(void) { return LGCA_B200_VERSION; }

int lgca_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lgca_b200_create(const lgca_b200_config* cfg, lgca_b200_lattice** out)
{
    if (!cfg || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->model < 0 || cfg->model > 3) return set_error(LGCA_B200_EINVAL, "invalid model %d", cfg->model);
    if (cfg->dim_x == 0 || cfg->dim_y == 0) return set_error(LGCA_B200_EINVAL, "lattice dimensions must be > 0");
    if (cfg->model != LGCA_B200_HPP && (cfg->dim_y & 1u))
        return set_error(LGCA_B200_EINVAL, "FHP models need an even dim_y (reference: src/lattice.cpp:141)");
    if (cfg->bf_dir != 0 && cfg->bf_dir != 'x' && cfg->bf_dir != 'y')
        return set_error(LGCA_B200_EINVAL, "bf_dir must be 'x', 'y' or 0");
    if (cfg->cg_radius && (cfg->dim_x % (2 * cfg->cg_radius) || cfg->dim_y % (2 * cfg->cg_radius) ||
                           cfg->dim_x < 4 * cfg->cg_radius))
        return set_error(LGCA_B200_EINVAL, "dims must be multiples of 2*cg_radius and dim_x >= 4*cg_radius "
                                           "(reference: src/lattice.cpp:152-153)");
    if (cfg->k_fuse < 0 || cfg->k_fuse > LGCA_MAX_K) return set_error(LGCA_B200_EINVAL, "k_fuse must be in [0, %d]", LGCA_MAX_K);

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return set_error(LGCA_B200_ENODEV, "no CUDA device: lgca_b200 has no CPU path");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return set_error(LGCA_B200_EINVAL, "device %d out of range", cfg->device);
    LGCA_CUDA_CHECK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    LGCA_CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return set_error(LGCA_B200_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device,
                         prop.major, prop.minor);

    lgca_b200_lattice* h = new lgca_b200_lattice();
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->nd  = num_dir_of(cfg->model);
    pthread_mutex_init(&h->snap_mutex, nullptr);
    h->snap_mutex_init = 1;
    // default fused depth: 6; FHP lattices too small to fill the machine (< 16 M sites, e.g. the 1400 x 700 pipe)
    // are launch-latency bound and run better with the shorter pipeline fill of 5
    h->k_fuse = cfg->k_fuse ? cfg->k_fuse
                            : ((cfg->model == LGCA_B200_HPP || (uint64_t)cfg->dim_x * (uint64_t)cfg->dim_y >= (1ull << 24)) ? 6 : 5);
    if (cfg->model != LGCA_B200_HPP && h->k_fuse > LGCA_MAX_K_FHP) h->k_fuse = LGCA_MAX_K_FHP;

    h->sm_count = prop.multiProcessorCount;
    const bool whole = (cfg->y_rows == 0 || cfg->y_rows == cfg->dim_y);
    if (!whole) {
        const uint32_t halo = (uint32_t)((std::max(h->k_fuse, 1) + 1) & ~1);
        int bad = 0;
        if ((uint64_t)cfg->y_begin + cfg->y_rows > cfg->dim_y) bad = set_error(LGCA_B200_EINVAL, "strip exceeds lattice");
        else if (cfg->cg_radius && (cfg->y_begin % (2 * cfg->cg_radius) || cfg->y_rows % (2 * cfg->cg_radius)))
            bad = set_error(LGCA_B200_EINVAL, "strip bounds must be multiples of 2*cg_radius");
        else if (cfg->model != LGCA_B200_HPP && ((cfg->y_begin | cfg->y_rows) & 1u))
            bad = set_error(LGCA_B200_EINVAL, "FHP strips must start on an even row and have an even height (hexagonal row parity)");
        else if (cfg->y_rows < halo)
            bad = set_error(LGCA_B200_EINVAL, "strip of %u rows is shorter than its halo of %u rows (k_fuse %d): lower k_fuse or use "
                                              "fewer strips", cfg->y_rows, halo, h->k_fuse);
        if (bad) { lgca_b200_destroy(h); return bad; }
    }
    Geom& g = h->g;
    g.dim_x  = cfg->dim_x;
    g.dim_y  = cfg->dim_y;
    g.nw     = (cfg->dim_x + 31) / 32;
    g.rem    = cfg->dim_x % 32;
    g.pitch  = (g.nw + 3) & ~3u;
    g.halo   = whole ? 0 : (uint32_t)((std::max(h->k_fuse, 1) + 1) & ~1); // even: keeps row parity
    g.rows   = (whole ? cfg->dim_y : cfg->y_rows) + 2 * g.halo;
    g.y0     = whole ? 0 : cfg->y_begin;
    g.wrap_y = whole ? 1 : 0;
    g.plane_stride = (uint64_t)g.rows * g.pitch;
    g.row_south = g.row_north = 0xFFFFFFFFu;
    for (uint32_t r = 0; r < g.rows; ++r) { // stored row r <-> global row (y0 - halo + r) mod dim_y
        const uint32_t gy = (uint32_t)(((uint64_t)g.y0 + g.dim_y - g.halo % g.dim_y + r) % g.dim_y);
        if (gy == 0) g.row_south = r;
        if (gy == g.dim_y - 1) g.row_north = r;
        if (r == 2 * g.halo + 1 && r + 2 * g.halo + 2 < g.rows) r = g.rows - 2 * g.halo - 2; // edges are near the ends
    }

    int rc = 0;
    const size_t pw = plane_words(h) * sizeof(uint32_t);
    for (int i = 0; i < 2 && !rc; ++i) rc = dev_alloc(h, (void**)&h->planes[i], pw * h->nd, true);
    if (!rc) rc = dev_alloc(h, (void**)&h->snap, pw * h->nd, true);
    h->base[0] = h->planes[0]; h->base[1] = h->planes[1]; h->base[2] = h->snap;
    if (!rc && g.halo) rc = dev_alloc(h, (void**)&h->snap_ghost, (size_t)g.pitch * sizeof(uint32_t) * h->nd, true);
    if (!rc) rc = dev_alloc(h, (void**)&h->ns, pw, true);
    if (!rc) rc = dev_alloc(h, (void**)&h->sl, pw, true);
    if (!rc) rc = dev_alloc(h, (void**)&h->ch, pw, true);
    if (!rc) rc = dev_alloc(h, (void**)&h->xedge, g.pitch * sizeof(uint32_t), true);
    if (!rc) rc = dev_alloc(h, (void**)&h->d_flags, 2 * sizeof(uint32_t), true);
    if (!rc) rc = dev_alloc(h, (void**)&h->d_scalars, 8 * sizeof(double), true);
    if (rc) { lgca_b200_destroy(h); return rc; }
    if (cudaHostAlloc((void**)&h->h_scalars, 8 * sizeof(double), cudaHostAllocDefault) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->s_post, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_post, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&h->ev_t0) != cudaSuccess || cudaEventCreate(&h->ev_t1) != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        lgca_b200_destroy(h);
        return set_cuda_error(e, "stream/event creation", __FILE__, __LINE__);
    }
    // the cudaMemset calls above ran in the default stream, which the handle's non-blocking streams do not wait for
    if (cudaDeviceSynchronize() != cudaSuccess) rc = set_cuda_error(cudaGetLastError(), "sync", __FILE__, __LINE__);
    if (!rc) rc = launch_build_xedge(h, h->s_compute);
    if (!rc && cudaStreamSynchronize(h->s_compute) != cudaSuccess) rc = set_cuda_error(cudaGetLastError(), "sync", __FILE__, __LINE__);
    if (rc) { lgca_b200_destroy(h); return rc; }
    // all-fluid, zero chirality, empty lattice is a valid starting point
    h->have_types = 1;
    *out = h;
    return 0;
}

int lgca_b200_destroy(lgca_b200_lattice* h)
{
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    lgca_b200_ring_disconnect(h);
    cudaFree(h->ring_flags);
    if (h->base[0] || h->base[1] || h->base[2]) {
        for (int i = 0; i < 3; ++i) cudaFree(h->base[i]); // the three plane sets, whatever their current roles
    } else {
        for (int i = 0; i < 2; ++i) cudaFree(h->planes[i]);
        cudaFree(h->snap);
    }
    for (int i = 0; i < 2; ++i) cudaFree(h->d_stage[i]);
    cudaFree(h->res_exch);
    free_mv_buffers(h);
    free_bf_buffers(h);
    cudaFree(h->snap_ghost); cudaFree(h->ns); cudaFree(h->sl); cudaFree(h->ch); cudaFree(h->xedge); cudaFree(h->d_flags);
    cudaFree(h->d_cell_density); cudaFree(h->d_cell_momentum); cudaFree(h->d_mean_density); cudaFree(h->d_mean_momentum);
    cudaFree(h->d_scalars); cudaFree(h->d_draws); cudaFree(h->d_draw_bytes);
    if (h->h_scalars) cudaFreeHost(h->h_scalars);
    if (h->h_draw_bytes) cudaFreeHost(h->h_draw_bytes);
    if (h->s_compute) cudaStreamDestroy(h->s_compute);
    if (h->s_post) cudaStreamDestroy(h->s_post);
    if (h->s_copy) cudaStreamDestroy(h->s_copy);
    for (int k = 0; k <= LGCA_MAX_K; ++k) { cudaFree(h->tile_fluid[k]); cudaFree(h->chain_done[k]); }
    if (h->snap_mutex_init) pthread_mutex_destroy(&h->snap_mutex);
    if (h->ev_snap) cudaEventDestroy(h->ev_snap);
    if (h->ev_post) cudaEventDestroy(h->ev_post);
    if (h->ev_t0) cudaEventDestroy(h->ev_t0);
    if (h->ev_t1) cudaEventDestroy(h->ev_t1);
    cudaGetLastError();
    delete h;
    return 0;
}

int lgca_b200_host_alloc(size_t bytes, void** out)
{
    if (!out) return set_error(LGCA_B200_EINVAL, "null argument");
    LGCA_CUDA_CHECK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}

int lgca_b200_host_free(void* p)
{
    if (p) LGCA_CUDA_CHECK(cudaFreeHost(p));
    return 0;
}

int lgca_b200_upload(lgca_b200_lattice* h, const uint8_t* state, const int32_t* cell_type, const uint8_t* rnd_bits)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int rc = ensure_stage(h);
    if (rc) return rc;
    const Geom& g = h->g;
    const uint32_t own = own_rows(h);
    cudaStream_t s = h->s_compute;
    int buf = 0;
    // the snapshot/post stream may still read the masks
    LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_post));

    if (cell_type) {
        const uint32_t rpc = (uint32_t)std::max<size_t>(1, h->stage_bytes / ((size_t)g.dim_x * 4));
        LGCA_CUDA_CHECK(cudaMemsetAsync(h->d_flags, 0, 2 * sizeof(uint32_t), s));
        for (uint32_t r0 = 0; r0 < own; r0 += rpc, buf ^= 1) {
            const uint32_t nr = std::min(rpc, own - r0);
            LGCA_CUDA_CHECK(cudaMemcpyAsync(h->d_stage[buf], cell_type + (size_t)r0 * g.dim_x, (size_t)nr * g.dim_x * 4,
                                            cudaMemcpyHostToDevice, s));
            if ((rc = launch_pack_cell_type(h, (const int32_t*)h->d_stage[buf], g.halo + r0, nr, s))) return rc;
        }
        uint32_t flags[2];
        LGCA_CUDA_CHECK(cudaMemcpyAsync(flags, h->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, s));
        LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
        h->has_ns = flags[0] != 0;
        h->has_sl = flags[1] != 0;
        memset(h->plan_valid, 0, sizeof(h->plan_valid));
        h->have_types = 1;
    }
    if (rnd_bits) {
        // rows per chunk such that the byte range of their bits fits the staging buffer
        const uint32_t rpc = (uint32_t)std::max<size_t>(1, (h->stage_bytes - 16) * 8 / g.dim_x);
        for (uint32_t r0 = 0; r0 < own; r0 += rpc, buf ^= 1) {
            const uint32_t nr = std::min(rpc, own - r0);
            const uint64_t first_bit = ((uint64_t)g.y0 + r0) * g.dim_x;
            const uint64_t last_bit  = first_bit + (uint64_t)nr * g.dim_x; // exclusive
            const uint64_t b0 = first_bit >> 3, b1 = (last_bit + 7) >> 3;
            LGCA_CUDA_CHECK(cudaMemcpyAsync(h->d_stage[buf], rnd_bits + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, s));
            if ((rc = launch_pack_rnd(h, (const uint8_t*)h->d_stage[buf], first_bit - (b0 << 3), g.halo + r0, nr, s))) return rc;
        }
        h->have_rnd = 1;
    }
    if (state) {
        if ((rc = unalias_snapshot(h))) return rc; // in-place write: the snapshot must not see it
        if ((rc = ring_order_inplace_write(h))) return rc; // ... and my last ghost-row push has read the old edge rows
        // two-stage pipeline over the staging buffers: the copy engine (s_copy) moves chunk c+1 over PCIe while the
        // SMs (compute stream) transpose chunk c into bit-planes
        if ((rc = ensure_copy_stream(h))) return rc;
        const uint32_t rpc = (uint32_t)std::max<size_t>(1, h->stage_bytes / g.dim_x);
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_stage_free[0], s)); // staging buffers are free once earlier work is done
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_stage_free[1], s));
        buf = 0;
        for (uint32_t r0 = 0; r0 < own; r0 += rpc, buf ^= 1) {
            const uint32_t nr = std::min(rpc, own - r0);
            LGCA_CUDA_CHECK(cudaStreamWaitEvent(h->s_copy, h->ev_stage_free[buf], 0));
            LGCA_CUDA_CHECK(cudaMemcpyAsync(h->d_stage[buf], state + (size_t)r0 * g.dim_x, (size_t)nr * g.dim_x,
                                            cudaMemcpyHostToDevice, h->s_copy));
            LGCA_CUDA_CHECK(cudaEventRecord(h->ev_stage_full[buf], h->s_copy));
            LGCA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_stage_full[buf], 0));
            if ((rc = launch_pack_state(h, (const uint8_t*)h->d_stage[buf], h->planes[h->cur], g.halo + r0, nr, s))) return rc;
            LGCA_CUDA_CHECK(cudaEventRecord(h->ev_stage_free[buf], s));
        }
        h->have_state = 1;
    }
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

int lgca_b200_download(lgca_b200_lattice* h, uint8_t* state)
{
    if (!h || !state) return set_error(LGCA_B200_EINVAL, "null argument");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int rc = ensure_stage(h);
    if (rc) return rc;
    const Geom& g = h->g;
    const uint32_t own = own_rows(h);
    cudaStream_t s = h->s_compute;
    const uint32_t rpc = (uint32_t)std::max<size_t>(1, h->stage_bytes / g.dim_x);
    // two-stage pipeline: the SMs unpack chunk c+1 while the copy engine sends chunk c over PCIe
    if ((rc = ensure_copy_stream(h))) return rc;
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_stage_free[0], h->s_copy));
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_stage_free[1], h->s_copy));
    int buf = 0;
    for (uint32_t r0 = 0; r0 < own; r0 += rpc, buf ^= 1) {
        const uint32_t nr = std::min(rpc, own - r0);
        LGCA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_stage_free[buf], 0));
        if ((rc = launch_unpack_state(h, h->planes[h->cur], (uint8_t*)h->d_stage[buf], g.halo + r0, nr, s))) return rc;
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_stage_full[buf], s));
        LGCA_CUDA_CHECK(cudaStreamWaitEvent(h->s_copy, h->ev_stage_full[buf], 0));
        LGCA_CUDA_CHECK(cudaMemcpyAsync(state + (size_t)r0 * g.dim_x, h->d_stage[buf], (size_t)nr * g.dim_x,
                                        cudaMemcpyDeviceToHost, h->s_copy));
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_stage_free[buf], h->s_copy));
    }
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_copy));
    return 0;
}

// chain_k != 0: the operation enqueued on the compute stream right before this call is a wave launch of k = chain_k
// (only event records in between): the first launch of this call may be chained to it
static int step_impl(lgca_b200_lattice* h, int n_steps, bool check_halo, int chain_k = 0)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    if (n_steps < 0) return set_error(LGCA_B200_EINVAL, "n_steps < 0");
    // a strip's ghost rows are only refreshed by the halo exchange, and the step kernels write owned rows only:
    // everything between two exchanges must fit ONE launch of the kernel that will actually run
    if (check_halo && h->g.halo && n_steps > 1 && n_steps > steps_per_launch(h, n_steps))
        return set_error(LGCA_B200_ESTATE, "this strip can advance at most %d step(s) between halo exchanges (k_fuse %d; the "
                                           "generic kernel advances one step per exchange)", steps_per_launch(h, h->k_fuse), h->k_fuse);
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    const bool simple = (h->cfg.flags & LGCA_B200_FLAG_SIMPLE_KERNEL) != 0;
    // lattices that fit on chip: ALL steps of the call in one launch of the SM-resident kernel
    if (n_steps >= 2 && !simple && resident_supported(h)) {
        int rc = launch_step_resident(h, h->planes[h->cur], h->planes[h->cur ^ 1], n_steps, h->s_compute);
        if (rc) return rc;
        h->cur ^= 1;
        if (h->snap_spare) { h->planes[h->cur ^ 1] = h->snap_spare; h->snap_spare = nullptr; }
        return 0;
    }
    int prev_k = chain_k; // k of the wave launch enqueued right before (0 = none)
    while (n_steps > 0) {
        int k = 1, rc;
        if (!simple) {
            k = std::min(h->k_fuse, n_steps);
            while (k > 1 && !wave_supported(h, k)) --k;
        }
        uint32_t* in  = h->planes[h->cur];
        uint32_t* out = h->planes[h->cur ^ 1];
        if (!simple && wave_supported(h, k)) { rc = launch_step_wave(h, in, out, k, h->s_compute, prev_k == k); prev_k = k; }
        else { k = 1; rc = launch_step_simple(h, in, out, h->s_compute); prev_k = 0; }
        if (rc) return rc;
        h->cur ^= 1;
        if (h->snap_spare) { // `in` is the zero-copy snapshot: the retired snapshot buffer takes its slot in the pair
            h->planes[h->cur ^ 1] = h->snap_spare;
            h->snap_spare = nullptr;
        }
        n_steps -= k;
    }
    return 0;
}

int lgca_b200_step(lgca_b200_lattice* h, int n_steps) { return step_impl(h, n_steps, true); }

} // extern "C"
int lgca_b200::step_one_launch(lgca_b200_lattice* h, int k, bool chain) { return step_impl(h, k, true, chain ? k : 0); }
extern "C" {

int lgca_b200_snapshot(lgca_b200_lattice* h)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    SnapLock lock(h);
    // do not overwrite the snapshot while the post stream still reads it
    LGCA_CUDA_CHECK(cudaStreamWaitEvent(h->s_compute, h->ev_post, 0));
    // native ring: the ghost rows of the live buffer are written by the neighbours; the coarse means of the top
    // coarse row read one of them, so the snapshot waits for the current epoch to have arrived
    if (h->ring_connected && h->ring_epoch) {
        int rc = ring_wait_current_epoch(h);
        if (rc) return rc;
    }
    if (h->g.halo) {
        // strips: the coarse means of the top coarse row read the first row beyond the owned rows; keep a side copy of
        // it, so that the post stream never touches ghost rows (the neighbours keep storing into them)
        const Geom& g = h->g;
        LGCA_CUDA_CHECK(cudaMemcpy2DAsync(h->snap_ghost, (size_t)g.pitch * sizeof(uint32_t),
                                          h->planes[h->cur] + (size_t)(g.rows - g.halo) * g.pitch,
                                          (size_t)g.plane_stride * sizeof(uint32_t), (size_t)g.pitch * sizeof(uint32_t), h->nd,
                                          cudaMemcpyDeviceToDevice, h->s_compute));
    }
    // zero-copy: the live buffer becomes the snapshot (see lgca_internal.h); nothing to do if it already is
    if (!h->snap_spare) {
        h->snap_spare = h->snap;
        h->snap = h->planes[h->cur];
    }
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_snap, h->s_compute));
    return 0;
}

int lgca_b200_post_process(lgca_b200_lattice* h, float* cell_density, float* cell_momentum, float* mean_density,
                           float* mean_momentum, int exact_order)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    const Geom& g = h->g;
    cudaStream_t s = h->s_post;
    const size_t cells = (size_t)g.dim_x * own_rows(h);
    int rc;
    SnapLock lock(h);
    LGCA_CUDA_CHECK(cudaStreamWaitEvent(s, h->ev_snap, 0));
    if (cell_density || cell_momentum) {
        if (h->cfg.flags & LGCA_B200_FLAG_NO_CELL_FIELDS)
            return set_error(LGCA_B200_EINVAL, "per-cell fields disabled by LGCA_B200_FLAG_NO_CELL_FIELDS");
        if (!h->d_cell_density) {
            if ((rc = dev_alloc(h, (void**)&h->d_cell_density, cells * sizeof(float), false))) return rc;
            if ((rc = dev_alloc(h, (void**)&h->d_cell_momentum, 2 * cells * sizeof(float), false))) return rc;
        }
        if ((rc = launch_cell_fields(h, h->snap, h->d_cell_density, h->d_cell_momentum, s))) return rc;
        if (cell_density)
            LGCA_CUDA_CHECK(cudaMemcpyAsync(cell_density, h->d_cell_density, cells * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (cell_momentum)
            LGCA_CUDA_CHECK(cudaMemcpyAsync(cell_momentum, h->d_cell_momentum, 2 * cells * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    if (mean_density || mean_momentum) {
        const uint32_t cg = h->cfg.cg_radius;
        if (!cg) return set_error(LGCA_B200_EINVAL, "coarse fields requested but cg_radius == 0");
        const size_t nc = (size_t)(g.dim_x / (2 * cg)) * (own_rows(h) / (2 * cg));
        if (!h->d_mean_density) {
            if ((rc = dev_alloc(h, (void**)&h->d_mean_density, std::max<size_t>(nc, 1) * sizeof(float), false))) return rc;
            if ((rc = dev_alloc(h, (void**)&h->d_mean_momentum, 2 * std::max<size_t>(nc, 1) * sizeof(float), false))) return rc;
        }
        if ((rc = launch_mean_fields(h, h->snap, h->g.halo ? h->snap_ghost : nullptr, h->d_mean_density, h->d_mean_momentum, exact_order, s))) return rc;
        if (mean_density)
            LGCA_CUDA_CHECK(cudaMemcpyAsync(mean_density, h->d_mean_density, nc * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (mean_momentum)
            LGCA_CUDA_CHECK(cudaMemcpyAsync(mean_momentum, h->d_mean_momentum, 2 * nc * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_post, s));
    lock.release(); // the stepping thread may snapshot again: its copy is ordered behind ev_post on the device
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

int lgca_b200_mean_velocity(lgca_b200_lattice* h, float out[2])
{
    if (!h || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    double s3[3];
    int rc = mean_velocity_sums(h, s3);
    if (rc) return rc;
    out[0] = (float)(s3[0] / s3[2]);
    out[1] = (float)(s3[1] / s3[2]);
    return 0;
}

int lgca_b200_count_particles(lgca_b200_lattice* h, uint64_t* out)
{
    if (!h || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->s_compute;
    unsigned long long* d = reinterpret_cast<unsigned long long*>(h->d_scalars + 4);
    int rc = launch_count_particles(h, h->planes[h->cur], d, s);
    if (rc) return rc;
    unsigned long long v = 0;
    LGCA_CUDA_CHECK(cudaMemcpyAsync(&v, d, sizeof(v), cudaMemcpyDeviceToHost, s));
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    *out = v;
    return 0;
}

// Exact body force (src/omp_lattice.cpp:254-346).  The reference is sequential: draw a cell, flip it if it is an
// eligible FLUID cell, stop after `forcing` flips.  Here the device GATHERs the byte states of a batch of drawn cells,
// the draws are REPLAYed in order on the host against those bytes (a cell changed by an earlier draw of the same
// batch is tracked), and the changed cells are APPLIED back.  The three stages are separate entry points so that a
// multi-GPU driver can combine the gathered bytes of all strips before the replay.

static int ensure_draw_buffers(lgca_b200_lattice* h, size_t n)
{
    if (n <= h->draw_cap) return 0;
    cudaFree(h->d_draws); cudaFree(h->d_draw_bytes);
    if (h->h_draw_bytes) cudaFreeHost(h->h_draw_bytes);
    h->d_draws = nullptr; h->d_draw_bytes = nullptr; h->h_draw_bytes = nullptr;
    h->draw_cap = 0;
    const size_t cap = std::max<size_t>(n, 1 << 16);
    // d_draws doubles as the scatter buffer: cap int32 cells followed by cap bytes
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_draws, cap * 5 + 16));
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_draw_bytes, cap));
    LGCA_CUDA_CHECK(cudaHostAlloc((void**)&h->h_draw_bytes, cap, cudaHostAllocDefault));
    h->draw_cap = cap;
    return 0;
}

int lgca_b200_body_force_gather(lgca_b200_lattice* h, const int32_t* cells, size_t n, uint8_t* bytes_out)
{
    if (!h || (n && (!cells || !bytes_out))) return set_error(LGCA_B200_EINVAL, "null argument");
    if (n == 0) return 0;
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int rc = ensure_draw_buffers(h, n);
    if (rc) return rc;
    cudaStream_t s = h->s_compute;
    LGCA_CUDA_CHECK(cudaMemcpyAsync(h->d_draws, cells, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if ((rc = launch_gather_cells(h, h->planes[h->cur], h->d_draws, n, h->d_draw_bytes, s))) return rc;
    LGCA_CUDA_CHECK(cudaMemcpyAsync(h->h_draw_bytes, h->d_draw_bytes, n, cudaMemcpyDeviceToHost, s));
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    memcpy(bytes_out, h->h_draw_bytes, n);
    return 0;
}

int lgca_b200_body_force_apply(lgca_b200_lattice* h, const int32_t* cells, const uint8_t* new_bytes, size_t n)
{
    if (!h || (n && (!cells || !new_bytes))) return set_error(LGCA_B200_EINVAL, "null argument");
    if (n == 0) return 0;
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int rc = ensure_draw_buffers(h, n);
    if (rc) return rc;
    cudaStream_t s = h->s_compute;
    std::vector<uint8_t> blob(n * 5);
    memcpy(blob.data(), cells, n * 4);
    memcpy(blob.data() + n * 4, new_bytes, n);
    LGCA_CUDA_CHECK(cudaMemcpyAsync(h->d_draws, blob.data(), blob.size(), cudaMemcpyHostToDevice, s));
    if ((rc = unalias_snapshot(h))) return rc; // in-place write: the snapshot must not see it
    if ((rc = ring_order_inplace_write(h))) return rc;
    if ((rc = launch_apply_flips(h, h->planes[h->cur], h->d_draws, n, s))) return rc;
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

// Host-only (no GPU): ordered replay of one batch.  bytes[i] = state byte of cells[i] with bit 7 set when the cell
// is not an eligible FLUID cell.  Stops after the draw that brings the reverted count to `forcing` (do-while of the
// reference: at least one draw is consumed even for forcing <= 0).  changed_* receive each changed cell once.
int lgca_b200_body_force_replay(int model, int bf_dir, int forcing, const int32_t* cells, const uint8_t* bytes, size_t n,
                                size_t* consumed, uint32_t* reverted, int32_t* changed_cells, uint8_t* changed_bytes,
                                size_t* n_changed)
{
    if ((n && (!cells || !bytes || !changed_cells || !changed_bytes)) || !consumed || !reverted || !n_changed)
        return set_error(LGCA_B200_EINVAL, "null argument");
    std::unordered_map<int32_t, uint8_t> touched;
    std::vector<int32_t> order; // first-touch order keeps the output deterministic
    size_t pos = 0;
    uint32_t rev = 0;
    const char bf = (char)bf_dir;
    bool done = false;
    while (pos < n && !done) {
        const int32_t cell = cells[pos];
        uint8_t b = bytes[pos];
        ++pos;
        if (!(b & 0x80)) {
            auto it = touched.find(cell);
            if (it != touched.end()) b = it->second;
            uint8_t w = b;
            if (model == LGCA_B200_HPP) { // src/omp_lattice.cpp:295-310
                if (bf == 'x' && !(b & 1) && (b & 4)) { w = (uint8_t)((w | 1) & ~4); ++rev; }
                else if (bf == 'y' && (b & 2) && !(b & 8)) { w = (uint8_t)((w | 8) & ~2); ++rev; }
            } else {                      // src/omp_lattice.cpp:313-338
                if (bf == 'x' && !(b & 1) && (b & 8)) { w = (uint8_t)((w | 1) & ~8); ++rev; }
                else if (bf == 'y') {
                    if ((b & 2) && !(b & 32)) { w = (uint8_t)((w | 32) & ~2); ++rev; }
                    if ((b & 4) && !(b & 16)) { w = (uint8_t)((w | 16) & ~4); ++rev; }
                }
            }
            if (w != b) {
                if (it == touched.end()) order.push_back(cell);
                touched[cell] = w;
            }
        }
        // do { ... } while (reverted_particles < forcing && ...): `unsigned int < int` compares as unsigned in the
        // reference (src/omp_lattice.cpp:264,346), so a negative forcing never stops on the count
        if (!(rev < (uint32_t)forcing)) done = true;
    }
    size_t k = 0;
    for (int32_t c : order) { changed_cells[k] = c; changed_bytes[k] = touched[c]; ++k; }
    *consumed = pos;
    *reverted = rev;
    *n_changed = k;
    return 0;
}

int lgca_b200_body_force(lgca_b200_lattice* h, int forcing, const int32_t* draws, size_t n_draws, size_t* consumed,
                         uint32_t* reverted)
{
    if (!h || (!draws && n_draws) || !consumed || !reverted) return set_error(LGCA_B200_EINVAL, "null argument");
    *consumed = 0;
    *reverted = 0;
    const Geom& g = h->g;
    const uint64_t num_cells = (uint64_t)g.dim_x * g.dim_y;
    if (num_cells > 0x7FFFFFFFull) return set_error(LGCA_B200_EINVAL, "body force needs < 2^31 cells (rand() range)");
    std::vector<int32_t> cells, ch_cells;
    std::vector<uint8_t> bytes, ch_bytes;
    size_t pos = 0;
    int64_t remaining = (int64_t)(uint32_t)forcing; // the reference's unsigned compare (negative forcing = "no limit")
    bool first = true;
    int rc;
    if (!g.halo && !(h->cfg.flags & LGCA_B200_FLAG_HOST_BODY_FORCE)) {
        // whole lattice: everything on the device (first occurrences, gains, prefix sum, stop rule, scatter)
        LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
        while (pos < n_draws && (first || remaining > 0)) {
            const size_t batch = std::min<size_t>(n_draws - pos, (size_t)1 << 22);
            size_t used = 0;
            uint32_t rev = 0;
            if ((rc = body_force_device(h, (uint32_t)remaining, first, draws + pos, batch, &used, &rev))) return rc;
            pos += used;
            remaining -= rev;
            *reverted += rev;
            first = false;
            if (used < batch) break; // stopped on the count
        }
        *consumed = pos;
        return 0;
    }
    while (pos < n_draws && (first || remaining > 0)) {
        size_t batch = (size_t)std::max<int64_t>(4096, std::min<int64_t>(1 << 20, remaining * 12));
        batch = std::min(batch, n_draws - pos);
        cells.resize(batch); bytes.resize(batch); ch_cells.resize(batch); ch_bytes.resize(batch);
        for (size_t i = 0; i < batch; ++i) cells[i] = (int32_t)((uint64_t)(uint32_t)draws[pos + i] % num_cells); // :269
        if ((rc = lgca_b200_body_force_gather(h, cells.data(), batch, bytes.data()))) return rc;
        size_t used = 0, nch = 0;
        uint32_t rev = 0;
        // a continuation batch must not re-run the reference's "at least one draw" rule
        if ((rc = lgca_b200_body_force_replay(h->cfg.model, h->cfg.bf_dir, (int)(uint32_t)(first ? remaining : std::max<int64_t>(remaining, 1)),
                                              cells.data(), bytes.data(), batch, &used, &rev, ch_cells.data(), ch_bytes.data(), &nch)))
            return rc;
        if ((rc = lgca_b200_body_force_apply(h, ch_cells.data(), ch_bytes.data(), nch))) return rc;
        pos += used;
        remaining -= rev;
        *reverted += rev;
        first = false;
    }
    *consumed = pos;
    return 0;
}

int lgca_b200_init_random_device(lgca_b200_lattice* h, uint64_t seed)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int rc = unalias_snapshot(h);
    if (!rc) rc = ring_order_inplace_write(h);
    if (!rc) rc = launch_init_random(h, h->planes[h->cur], seed, h->s_compute);
    if (rc) return rc;
    LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
    h->have_state = h->have_rnd = 1;
    return 0;
}

int lgca_b200_apply_bc_device(lgca_b200_lattice* h, const char* bc)
{
    if (!h || !bc) return set_error(LGCA_B200_EINVAL, "null argument");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int kind;
    if (!strcmp(bc, "periodic")) kind = 0;
    else if (!strcmp(bc, "pipe")) kind = 1;
    else if (!strcmp(bc, "karman")) kind = 2;
    else if (!strcmp(bc, "reflecting_back")) kind = 3;
    else if (!strcmp(bc, "reflecting_forward")) kind = 4;
    else return set_error(LGCA_B200_EINVAL, "unknown bc '%s'", bc);
    LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_post));
    int rc = launch_paint_bc(h, kind, h->s_compute);
    if (rc) return rc;
    LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
    h->has_ns = (kind >= 1 && kind <= 3);
    h->has_sl = (kind == 4);
    memset(h->plan_valid, 0, sizeof(h->plan_valid));
    h->have_types = 1;
    return 0;
}

int lgca_b200_sync(lgca_b200_lattice* h)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
    LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_post));
    if (h->s_ring) LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_ring));
    return 0;
}

void* lgca_b200_compute_stream(lgca_b200_lattice* h) { return h ? (void*)h->s_compute : nullptr; }

int lgca_b200_timed_steps(lgca_b200_lattice* h, int n_steps, float* elapsed_ms)
{
    if (!h || !elapsed_ms) return set_error(LGCA_B200_EINVAL, "null argument");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_t0, h->s_compute));
    int rc = lgca_b200_step(h, n_steps);
    if (rc) return rc;
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_t1, h->s_compute));
    LGCA_CUDA_CHECK(cudaEventSynchronize(h->ev_t1));
    LGCA_CUDA_CHECK(cudaEventElapsedTime(elapsed_ms, h->ev_t0, h->ev_t1));
    return 0;
}

int lgca_b200_timed_kernel(lgca_b200_lattice* h, int launches, float* ms_per_launch)
{
    if (!h || !ms_per_launch || launches <= 0) return set_error(LGCA_B200_EINVAL, "bad argument");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    const int k = steps_per_launch(h, h->k_fuse);
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_t0, h->s_compute));
    int rc = step_impl(h, k * launches, false);
    if (rc) return rc;
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_t1, h->s_compute));
    LGCA_CUDA_CHECK(cudaEventSynchronize(h->ev_t1));
    float ms = 0;
    LGCA_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
    *ms_per_launch = ms / launches;
    return 0;
}

int lgca_b200_launch_count(lgca_b200_lattice* h, uint64_t* out)
{
    if (!h || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    *out = h->launches;
    return 0;
}

int lgca_b200_get_info(lgca_b200_lattice* h, lgca_b200_info* out)
{
    if (!h || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    memset(out, 0, sizeof(*out));
    out->dim_x = h->g.dim_x; out->dim_y = h->g.dim_y; out->y_begin = h->g.y0; out->y_rows = own_rows(h);
    out->words_per_row = h->g.pitch; out->num_planes = (uint32_t)h->nd;
    out->has_no_slip = (uint32_t)h->has_ns; out->has_slip = (uint32_t)h->has_sl;
    out->k_fuse = h->k_fuse;
    // information-carrying mask planes read per step: chirality (FHP), no-slip, slip
    out->bytes_per_site_step_x8 = 2u * h->nd + (h->cfg.model != LGCA_B200_HPP ? 1u : 0u) + (h->has_ns ? 1u : 0u) +
                                  (h->has_sl ? 1u : 0u);
    out->device_bytes = h->device_bytes;
    return 0;
}

int lgca_b200_halo_rows(lgca_b200_lattice* h, uint32_t* rows)
{
    if (!h || !rows) return set_error(LGCA_B200_EINVAL, "null argument");
    *rows = h->g.halo;
    return 0;
}

int lgca_b200_steps_per_exchange(lgca_b200_lattice* h, int* steps)
{
    if (!h || !steps) return set_error(LGCA_B200_EINVAL, "null argument");
    *steps = h->g.halo ? std::min(steps_per_launch(h, h->k_fuse), (int)h->g.halo) : h->k_fuse;
    return 0;
}

static int halo_planes(const lgca_b200_lattice* h, int what) { return what == LGCA_B200_HALO_STATE ? h->nd : 3; }

int lgca_b200_halo_bytes(lgca_b200_lattice* h, int what, size_t* bytes_per_side)
{
    if (!h || !bytes_per_side) return set_error(LGCA_B200_EINVAL, "null argument");
    if (what != LGCA_B200_HALO_STATE && what != LGCA_B200_HALO_MASKS) return set_error(LGCA_B200_EINVAL, "bad halo kind");
    *bytes_per_side = (size_t)h->g.halo * h->g.pitch * sizeof(uint32_t) * halo_planes(h, what);
    return 0;
}

// packed halo layout: [plane][halo row][pitch words]
static int halo_copy(lgca_b200_lattice* h, int what, void* packed, uint32_t first_row, bool to_packed)
{
    const Geom& g = h->g;
    const size_t row_bytes = (size_t)g.pitch * sizeof(uint32_t);
    const int np = halo_planes(h, what);
    for (int d = 0; d < np; ++d) {
        uint32_t* base;
        if (what == LGCA_B200_HALO_STATE) base = h->planes[h->cur] + (size_t)d * g.plane_stride;
        else base = d == 0 ? h->ns : (d == 1 ? h->sl : h->ch);
        uint32_t* pl = base + (size_t)first_row * g.pitch;
        uint8_t*  pk = (uint8_t*)packed + (size_t)d * g.halo * row_bytes;
        if (to_packed) LGCA_CUDA_CHECK(cudaMemcpyAsync(pk, pl, g.halo * row_bytes, cudaMemcpyDeviceToDevice, h->s_compute));
        else LGCA_CUDA_CHECK(cudaMemcpyAsync(pl, pk, g.halo * row_bytes, cudaMemcpyDeviceToDevice, h->s_compute));
    }
    return 0;
}

int lgca_b200_halo_export(lgca_b200_lattice* h, int what, void* dev_top_rows, void* dev_bottom_rows)
{
    if (!h || !dev_top_rows || !dev_bottom_rows) return set_error(LGCA_B200_EINVAL, "null argument");
    if (!h->g.halo) return set_error(LGCA_B200_ESTATE, "handle owns the whole lattice: no halo");
    if (what != LGCA_B200_HALO_STATE && what != LGCA_B200_HALO_MASKS) return set_error(LGCA_B200_EINVAL, "bad halo kind");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    const Geom& g = h->g;
    // "top" = the strip's highest own rows (they become the upper neighbour's lower halo)
    int rc = halo_copy(h, what, dev_top_rows, g.rows - 2 * g.halo, true);
    if (!rc) rc = halo_copy(h, what, dev_bottom_rows, g.halo, true);
    return rc;
}

int lgca_b200_halo_import(lgca_b200_lattice* h, int what, const void* dev_from_upper, const void* dev_from_lower)
{
    if (!h || !dev_from_upper || !dev_from_lower) return set_error(LGCA_B200_EINVAL, "null argument");
    if (!h->g.halo) return set_error(LGCA_B200_ESTATE, "handle owns the whole lattice: no halo");
    if (what != LGCA_B200_HALO_STATE && what != LGCA_B200_HALO_MASKS) return set_error(LGCA_B200_EINVAL, "bad halo kind");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    const Geom& g = h->g;
    // the upper neighbour's bottom rows fill the halo above the strip, and vice versa
    int rc = halo_copy(h, what, const_cast<void*>(dev_from_upper), g.rows - g.halo, false);
    if (!rc) rc = halo_copy(h, what, const_cast<void*>(dev_from_lower), 0, false);
    if (what == LGCA_B200_HALO_MASKS) memset(h->plan_valid, 0, sizeof(h->plan_valid)); // ghost-row masks changed: per-tile flags are stale
    return rc;
}

// Wall-kind flags select the kernel variant; with row strips every rank must use the union of all
// strips' flags (a strip without walls of its own may import wall cells in its halo rows).
int lgca_b200_get_wall_flags(lgca_b200_lattice* h, uint32_t* has_no_slip, uint32_t* has_slip)
{
    if (!h || !has_no_slip || !has_slip) return set_error(LGCA_B200_EINVAL, "null argument");
    *has_no_slip = (uint32_t)h->has_ns;
    *has_slip = (uint32_t)h->has_sl;
    return 0;
}

int lgca_b200_set_wall_flags(lgca_b200_lattice* h, uint32_t has_no_slip, uint32_t has_slip)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    h->has_ns = has_no_slip != 0;
    h->has_sl = has_slip != 0;
    memset(h->plan_valid, 0, sizeof(h->plan_valid));
    return 0;
}

} // extern "C"
