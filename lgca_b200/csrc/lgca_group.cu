// lgca_b200_group_*: one lattice on several GPUs of one box, driven by ONE host process (include/lgca_b200.h).
//
// The reference has a single shared-memory lattice behind Lattice<Model> (src/lattice.h:187-211); this is what lets the
// drop-in C++ backend (host/b200_lattice.cpp) and the headless apps (--gpus N) spread that lattice over N devices
// without the caller seeing strips: all host arrays are GLOBAL reference-layout arrays.  Host orchestration only --
// every operation is a loop over the strips' single-handle calls:
//   * rows are cut into contiguous strips (heights multiples of 2*cg_radius, even), one handle per device;
//   * the strips are wired into the native halo ring through same-process peer pointers (lgca_ring.cu);
//   * stepping is issued BLOCK-MAJOR (for every block of fused steps: every strip) so that no strip's launch queue can
//     fill up with kernels that wait for a neighbour whose work has not been enqueued yet;
//   * the body force gathers the drawn cells on every strip, combines (minimum), replays once on the host, applies on
//     every strip and republishes the edge rows;
//   * snapshots rotate plane sets in lockstep on all strips (zero copy).
// With one device the group is a plain whole-lattice handle.
#include <string.h>

#include <algorithm>
#include <vector>

#include "lgca_internal.h"

struct lgca_b200_group {
    lgca_b200_config                cfg;
    std::vector<lgca_b200_lattice*> strips;
    std::vector<uint32_t>           y0, rows;
    bool                            ring_connected = false;
    bool                            published = false; // edge rows have been published since the last in-place write
};

using namespace lgca_b200;

namespace {

inline size_t n_strips(const lgca_b200_group* g) { return g->strips.size(); }
inline bool   multi(const lgca_b200_group* g) { return g->strips.size() > 1; }

// static mask planes of the ghost rows: packed export on every strip, peer copies, import (set-up path, synchronous)
int exchange_masks(lgca_b200_group* g)
{
    const size_t n = n_strips(g);
    std::vector<void*> top(n, nullptr), bottom(n, nullptr), from_upper(n, nullptr), from_lower(n, nullptr);
    int rc = 0;
    size_t bytes = 0;
    auto cleanup = [&]() {
        for (size_t i = 0; i < n; ++i) {
            cudaSetDevice(g->strips[i]->cfg.device);
            cudaFree(top[i]); cudaFree(bottom[i]); cudaFree(from_upper[i]); cudaFree(from_lower[i]);
        }
    };
    for (size_t i = 0; i < n && !rc; ++i) {
        lgca_b200_lattice* h = g->strips[i];
        if ((rc = lgca_b200_halo_bytes(h, LGCA_B200_HALO_MASKS, &bytes))) break;
        if (cudaSetDevice(h->cfg.device) != cudaSuccess || cudaMalloc(&top[i], bytes) != cudaSuccess ||
            cudaMalloc(&bottom[i], bytes) != cudaSuccess || cudaMalloc(&from_upper[i], bytes) != cudaSuccess ||
            cudaMalloc(&from_lower[i], bytes) != cudaSuccess) {
            rc = set_cuda_error(cudaGetLastError(), "mask halo buffers", __FILE__, __LINE__);
            break;
        }
        if ((rc = lgca_b200_halo_export(h, LGCA_B200_HALO_MASKS, top[i], bottom[i]))) break;
        if ((rc = lgca_b200_sync(h))) break;
    }
    for (size_t i = 0; i < n && !rc; ++i) {
        const size_t up = (i + 1) % n, lo = (i + n - 1) % n;
        // my top rows are the upper neighbour's lower ghost rows, my bottom rows the lower neighbour's upper ones
        if (cudaMemcpyPeer(from_lower[up], g->strips[up]->cfg.device, top[i], g->strips[i]->cfg.device, bytes) != cudaSuccess ||
            cudaMemcpyPeer(from_upper[lo], g->strips[lo]->cfg.device, bottom[i], g->strips[i]->cfg.device, bytes) != cudaSuccess)
            rc = set_cuda_error(cudaGetLastError(), "cudaMemcpyPeer (mask halo)", __FILE__, __LINE__);
    }
    // cudaMemcpyPeer may return before the copy has finished and is not ordered against the handles' non-blocking
    // streams: drain every device before the imports read the buffers
    for (size_t i = 0; i < n && !rc; ++i)
        if (cudaSetDevice(g->strips[i]->cfg.device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
            rc = set_cuda_error(cudaGetLastError(), "cudaDeviceSynchronize (mask halo)", __FILE__, __LINE__);
    for (size_t i = 0; i < n && !rc; ++i) {
        if ((rc = lgca_b200_halo_import(g->strips[i], LGCA_B200_HALO_MASKS, from_upper[i], from_lower[i]))) break;
        rc = lgca_b200_sync(g->strips[i]);
    }
    cleanup();
    return rc;
}

int unify_wall_flags(lgca_b200_group* g)
{
    uint32_t ns = 0, sl = 0;
    for (lgca_b200_lattice* h : g->strips) {
        uint32_t a = 0, b = 0;
        int rc = lgca_b200_get_wall_flags(h, &a, &b);
        if (rc) return rc;
        ns |= a; sl |= b;
    }
    for (lgca_b200_lattice* h : g->strips) {
        int rc = lgca_b200_set_wall_flags(h, ns, sl);
        if (rc) return rc;
    }
    return 0;
}

int connect_ring(lgca_b200_group* g)
{
    if (g->ring_connected || !multi(g)) return 0;
    const size_t n = n_strips(g);
    size_t bytes = 0;
    int rc = lgca_b200_ring_descriptor_bytes(&bytes);
    if (rc) return rc;
    std::vector<std::vector<uint8_t>> desc(n, std::vector<uint8_t>(bytes));
    for (size_t i = 0; i < n; ++i)
        if ((rc = lgca_b200_ring_export(g->strips[i], desc[i].data(), bytes))) return rc;
    for (size_t i = 0; i < n; ++i)
        if ((rc = lgca_b200_ring_connect(g->strips[i], desc[(i + n - 1) % n].data(), desc[(i + 1) % n].data()))) return rc;
    g->ring_connected = true;
    return 0;
}

// collective publish of the edge rows (first start or after an in-place write)
int publish(lgca_b200_group* g)
{
    if (!multi(g)) return 0;
    int rc = connect_ring(g);
    if (rc) return rc;
    for (lgca_b200_lattice* h : g->strips)
        if ((rc = lgca_b200_ring_republish(h))) return rc;
    g->published = true;
    return 0;
}

int ensure_published(lgca_b200_group* g) { return (multi(g) && !g->published) ? publish(g) : 0; }

} // namespace

extern "C" {

int lgca_b200_group_create(const lgca_b200_config* cfg, int n_gpus, const int* dev_ids, lgca_b200_group** out)
{
    if (!cfg || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    *out = nullptr;
    if (n_gpus < 1) return set_error(LGCA_B200_EINVAL, "n_gpus must be >= 1");
    const int ndev = lgca_b200_device_count();
    if (ndev == 0) return set_error(LGCA_B200_ENODEV, "no CUDA device: lgca_b200 has no CPU path");
    if (!dev_ids && n_gpus > ndev) return set_error(LGCA_B200_EINVAL, "%d GPUs requested, %d visible", n_gpus, ndev);
    for (int i = 0; dev_ids && i < n_gpus; ++i) // an ordinal may repeat: several strips on one device (testing)
        if (dev_ids[i] < 0 || dev_ids[i] >= ndev) return set_error(LGCA_B200_EINVAL, "device %d out of range (%d visible)", dev_ids[i], ndev);
    // strips: contiguous, heights multiples of `unit` rows (coarse cells stay strip-local, hex row parity stays global)
    const uint32_t unit = std::max<uint32_t>(2u * cfg->cg_radius, 2u);
    if (n_gpus > 1 && (cfg->dim_y % unit || cfg->dim_y / unit < (uint32_t)n_gpus))
        return set_error(LGCA_B200_EINVAL, "dim_y = %u cannot be cut into %d strips of multiples of %u rows", cfg->dim_y, n_gpus, unit);
    lgca_b200_group* g = new lgca_b200_group();
    g->cfg = *cfg;
    const uint32_t units = cfg->dim_y / unit, base = units / (uint32_t)n_gpus, extra = units % (uint32_t)n_gpus;
    uint32_t y = 0;
    for (int i = 0; i < n_gpus; ++i) {
        lgca_b200_config c = *cfg;
        c.device = dev_ids ? dev_ids[i] : i;
        if (n_gpus == 1) { c.y_begin = 0; c.y_rows = 0; }
        else {
            c.y_begin = y;
            c.y_rows = (base + ((uint32_t)i < extra ? 1u : 0u)) * unit;
            y += c.y_rows;
        }
        lgca_b200_lattice* h = nullptr;
        const int rc = lgca_b200_create(&c, &h);
        if (rc) { lgca_b200_group_destroy(g); return rc; }
        g->strips.push_back(h);
        g->y0.push_back(n_gpus == 1 ? 0u : c.y_begin);
        g->rows.push_back(n_gpus == 1 ? cfg->dim_y : c.y_rows);
    }
    *out = g;
    return 0;
}

int lgca_b200_group_destroy(lgca_b200_group* g)
{
    if (!g) return 0;
    // drain every device first: a strip's ring kernels may still wait for a neighbour
    for (lgca_b200_lattice* h : g->strips) lgca_b200_sync(h);
    for (lgca_b200_lattice* h : g->strips) lgca_b200_ring_disconnect(h);
    for (lgca_b200_lattice* h : g->strips) lgca_b200_destroy(h);
    delete g;
    return 0;
}

int lgca_b200_group_size(lgca_b200_group* g, int* n_gpus)
{
    if (!g || !n_gpus) return set_error(LGCA_B200_EINVAL, "null argument");
    *n_gpus = (int)n_strips(g);
    return 0;
}

int lgca_b200_group_strip(lgca_b200_group* g, int i, lgca_b200_lattice** h)
{
    if (!g || !h || i < 0 || i >= (int)n_strips(g)) return set_error(LGCA_B200_EINVAL, "bad argument");
    *h = g->strips[(size_t)i];
    return 0;
}

int lgca_b200_group_upload(lgca_b200_group* g, const uint8_t* state, const int32_t* cell_type, const uint8_t* rnd_bits)
{
    if (!g) return set_error(LGCA_B200_EINVAL, "null group");
    const size_t dx = g->cfg.dim_x;
    int rc;
    for (size_t i = 0; i < n_strips(g); ++i) {
        const size_t off = (size_t)g->y0[i] * dx;
        if ((rc = lgca_b200_upload(g->strips[i], state ? state + off : nullptr, cell_type ? cell_type + off : nullptr, rnd_bits)))
            return rc;
    }
    if (!multi(g)) return 0;
    if (cell_type && (rc = unify_wall_flags(g))) return rc;
    if ((cell_type || rnd_bits) && (rc = exchange_masks(g))) return rc;
    if (state || !g->published) { g->published = false; if ((rc = publish(g))) return rc; }
    return 0;
}

int lgca_b200_group_download(lgca_b200_group* g, uint8_t* state)
{
    if (!g || !state) return set_error(LGCA_B200_EINVAL, "null argument");
    for (size_t i = 0; i < n_strips(g); ++i) {
        const int rc = lgca_b200_download(g->strips[i], state + (size_t)g->y0[i] * g->cfg.dim_x);
        if (rc) return rc;
    }
    return 0;
}

int lgca_b200_group_step(lgca_b200_group* g, int n_steps)
{
    if (!g) return set_error(LGCA_B200_EINVAL, "null group");
    if (n_steps < 0) return set_error(LGCA_B200_EINVAL, "n_steps < 0");
    if (!multi(g)) return lgca_b200_step(g->strips[0], n_steps);
    int rc = ensure_published(g);
    if (rc) return rc;
    int block = 1 << 30;
    for (lgca_b200_lattice* h : g->strips) {
        int b = 1;
        if ((rc = lgca_b200_steps_per_exchange(h, &b))) return rc;
        block = std::min(block, b);
    }
    bool first = true;
    while (n_steps > 0) { // block-major: see the top of the file
        const int k = std::min(block, n_steps);
        // nothing else touches a strip's compute stream between two blocks of this loop: launches may be chained
        for (lgca_b200_lattice* h : g->strips)
            if ((rc = ring_step_blocks(h, k, !first))) return rc;
        first = false;
        n_steps -= k;
    }
    return 0;
}

int lgca_b200_group_snapshot(lgca_b200_group* g)
{
    if (!g) return set_error(LGCA_B200_EINVAL, "null group");
    int rc = ensure_published(g);
    if (rc) return rc;
    for (lgca_b200_lattice* h : g->strips)
        if ((rc = lgca_b200_snapshot(h))) return rc;
    return 0;
}

int lgca_b200_group_post_process(lgca_b200_group* g, float* cell_density, float* cell_momentum, float* mean_density,
                                 float* mean_momentum, int exact_order)
{
    if (!g) return set_error(LGCA_B200_EINVAL, "null group");
    const size_t dx = g->cfg.dim_x, cg = g->cfg.cg_radius;
    const size_t cdx = cg ? dx / (2 * cg) : 0;
    for (size_t i = 0; i < n_strips(g); ++i) {
        const size_t cell0 = (size_t)g->y0[i] * dx;
        const size_t coarse0 = cg ? (size_t)(g->y0[i] / (2 * cg)) * cdx : 0;
        const int rc = lgca_b200_post_process(g->strips[i], cell_density ? cell_density + cell0 : nullptr,
                                              cell_momentum ? cell_momentum + 2 * cell0 : nullptr,
                                              mean_density ? mean_density + coarse0 : nullptr,
                                              mean_momentum ? mean_momentum + 2 * coarse0 : nullptr, exact_order);
        if (rc) return rc;
    }
    return 0;
}

int lgca_b200_group_mean_velocity(lgca_b200_group* g, float out[2])
{
    if (!g || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    double sx = 0, sy = 0, cnt = 0;
    for (lgca_b200_lattice* h : g->strips) {
        double s3[3];
        const int rc = mean_velocity_sums(h, s3);
        if (rc) return rc;
        sx += s3[0]; sy += s3[1]; cnt += s3[2];
    }
    out[0] = (float)(sx / cnt);
    out[1] = (float)(sy / cnt);
    return 0;
}

int lgca_b200_group_mean_velocity_exact(lgca_b200_group* g, float out[2])
{
    if (!g || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    float    sums[2] = {0.0f, 0.0f};
    uint64_t cnt = 0;
    for (lgca_b200_lattice* h : g->strips) { // strips are stored in y order = cell order
        const int rc = lgca_b200_mean_velocity_exact(h, sums, &cnt);
        if (rc) return rc;
    }
    out[0] = sums[0] / (float)cnt; // src/omp_lattice.cpp:553-554
    out[1] = sums[1] / (float)cnt;
    return 0;
}

// Exact body force over the strips (src/omp_lattice.cpp:254-346): the same batches as lgca_b200_body_force, with the
// gather on every strip (0x80 = "not mine / not eligible"), an element-wise minimum, ONE ordered host replay, the apply on
// every strip, and a republish of the edge rows when anything changed.
int lgca_b200_group_body_force(lgca_b200_group* g, int forcing, const int32_t* draws, size_t n_draws, size_t* consumed,
                               uint32_t* reverted)
{
    if (!g || (!draws && n_draws) || !consumed || !reverted) return set_error(LGCA_B200_EINVAL, "null argument");
    if (!multi(g)) return lgca_b200_body_force(g->strips[0], forcing, draws, n_draws, consumed, reverted);
    *consumed = 0;
    *reverted = 0;
    const uint64_t num_cells = (uint64_t)g->cfg.dim_x * g->cfg.dim_y;
    if (num_cells > 0x7FFFFFFFull) return set_error(LGCA_B200_EINVAL, "body force needs < 2^31 cells (rand() range)");
    int rc = ensure_published(g);
    if (rc) return rc;
    std::vector<int32_t> cells, ch_cells;
    std::vector<uint8_t> bytes, part, ch_bytes;
    size_t pos = 0;
    int64_t remaining = (int64_t)(uint32_t)forcing;
    bool first = true, changed = false;
    if (!(g->cfg.flags & LGCA_B200_FLAG_HOST_BODY_FORCE)) {
        // device path: classification on every strip, gains summed on the first strip (peer copies), prefix sum + stop rule
        // there, scatter of the own reverts on every strip (csrc/lgca_bodyforce.cu)
        lgca_b200_lattice* h0 = g->strips[0];
        while (pos < n_draws && (first || remaining > 0)) {
            const size_t batch = std::min<size_t>(n_draws - pos, (size_t)1 << 22);
            for (lgca_b200_lattice* h : g->strips)
                if ((rc = body_force_classify(h, draws + pos, batch))) return rc;
            for (size_t i = 1; i < n_strips(g); ++i)
                if ((rc = body_force_combine(h0, g->strips[i], batch))) return rc;
            size_t used = 0;
            uint32_t rev = 0;
            if ((rc = body_force_cutoff(h0, (uint32_t)remaining, first, batch, &used, &rev))) return rc;
            if (rev) {
                // every strip gets the call: the copy-on-write of the snapshot must happen on all strips or on none
                for (lgca_b200_lattice* h : g->strips)
                    if ((rc = body_force_apply_cut(h, batch, used))) return rc;
                changed = true;
            }
            pos += used;
            remaining -= rev;
            *reverted += rev;
            first = false;
            if (used < batch) break;
        }
        *consumed = pos;
        if (changed) { g->published = false; if ((rc = publish(g))) return rc; }
        return 0;
    }
    while (pos < n_draws && (first || remaining > 0)) {
        size_t batch = (size_t)std::max<int64_t>(4096, std::min<int64_t>(1 << 20, remaining * 12));
        batch = std::min(batch, n_draws - pos);
        cells.resize(batch); bytes.assign(batch, 0xFF); part.resize(batch); ch_cells.resize(batch); ch_bytes.resize(batch);
        for (size_t i = 0; i < batch; ++i) cells[i] = (int32_t)((uint64_t)(uint32_t)draws[pos + i] % num_cells);
        for (lgca_b200_lattice* h : g->strips) {
            if ((rc = lgca_b200_body_force_gather(h, cells.data(), batch, part.data()))) return rc;
            for (size_t i = 0; i < batch; ++i) bytes[i] = std::min(bytes[i], part[i]);
        }
        size_t used = 0, nch = 0;
        uint32_t rev = 0;
        if ((rc = lgca_b200_body_force_replay(g->cfg.model, g->cfg.bf_dir, (int)(uint32_t)(first ? remaining : std::max<int64_t>(remaining, 1)),
                                              cells.data(), bytes.data(), batch, &used, &rev, ch_cells.data(), ch_bytes.data(), &nch)))
            return rc;
        if (nch) {
            // every strip gets the call (cells outside a strip are ignored there): the copy-on-write of the snapshot
            // must happen on all strips or on none, or their buffer rotation would fall out of step
            for (lgca_b200_lattice* h : g->strips)
                if ((rc = lgca_b200_body_force_apply(h, ch_cells.data(), ch_bytes.data(), nch))) return rc;
            changed = true;
        }
        pos += used;
        remaining -= rev;
        *reverted += rev;
        first = false;
    }
    *consumed = pos;
    if (changed) { g->published = false; if ((rc = publish(g))) return rc; }
    return 0;
}

int lgca_b200_group_count_particles(lgca_b200_group* g, uint64_t* out)
{
    if (!g || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    uint64_t total = 0;
    for (lgca_b200_lattice* h : g->strips) {
        uint64_t v = 0;
        const int rc = lgca_b200_count_particles(h, &v);
        if (rc) return rc;
        total += v;
    }
    *out = total;
    return 0;
}

int lgca_b200_group_init_random_device(lgca_b200_group* g, uint64_t seed)
{
    if (!g) return set_error(LGCA_B200_EINVAL, "null group");
    for (lgca_b200_lattice* h : g->strips) { // the hash is keyed on the GLOBAL cell: ghost rows come out right by themselves
        const int rc = lgca_b200_init_random_device(h, seed);
        if (rc) return rc;
    }
    g->published = false;
    return publish(g);
}

int lgca_b200_group_apply_bc_device(lgca_b200_group* g, const char* bc)
{
    if (!g) return set_error(LGCA_B200_EINVAL, "null group");
    for (lgca_b200_lattice* h : g->strips) { // painted per global row, ghost rows included
        const int rc = lgca_b200_apply_bc_device(h, bc);
        if (rc) return rc;
    }
    return 0; // every strip set the same wall flags from the BC kind
}

int lgca_b200_group_sync(lgca_b200_group* g)
{
    if (!g) return set_error(LGCA_B200_EINVAL, "null group");
    for (lgca_b200_lattice* h : g->strips) {
        const int rc = lgca_b200_sync(h);
        if (rc) return rc;
    }
    return 0;
}

int lgca_b200_group_timed_steps(lgca_b200_group* g, int n_steps, float* elapsed_ms)
{
    if (!g || !elapsed_ms) return set_error(LGCA_B200_EINVAL, "null argument");
    if (!multi(g)) return lgca_b200_timed_steps(g->strips[0], n_steps, elapsed_ms);
    int rc = ensure_published(g);
    if (rc) return rc;
    for (lgca_b200_lattice* h : g->strips) {
        LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_t0, h->s_compute));
    }
    if ((rc = lgca_b200_group_step(g, n_steps))) return rc;
    for (lgca_b200_lattice* h : g->strips) {
        LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_t1, h->s_compute));
    }
    float worst = 0;
    for (lgca_b200_lattice* h : g->strips) {
        LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
        LGCA_CUDA_CHECK(cudaEventSynchronize(h->ev_t1));
        float ms = 0;
        LGCA_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev_t0, h->ev_t1));
        worst = std::max(worst, ms);
    }
    *elapsed_ms = worst;
    return 0;
}

int lgca_b200_group_launch_count(lgca_b200_group* g, uint64_t* out)
{
    if (!g || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    uint64_t total = 0;
    for (lgca_b200_lattice* h : g->strips) total += h->launches;
    *out = total;
    return 0;
}

int lgca_b200_group_get_info(lgca_b200_group* g, lgca_b200_info* out)
{
    if (!g || !out) return set_error(LGCA_B200_EINVAL, "null argument");
    int rc = lgca_b200_get_info(g->strips[0], out);
    if (rc) return rc;
    out->y_begin = 0;
    out->y_rows = g->cfg.dim_y;
    uint64_t bytes = 0;
    int block = 1 << 30;
    for (lgca_b200_lattice* h : g->strips) {
        lgca_b200_info i;
        if ((rc = lgca_b200_get_info(h, &i))) return rc;
        bytes += i.device_bytes;
        out->has_no_slip |= i.has_no_slip;
        out->has_slip |= i.has_slip;
        int b = 1;
        if ((rc = lgca_b200_steps_per_exchange(h, &b))) return rc;
        block = std::min(block, b);
    }
    out->device_bytes = bytes;
    out->k_fuse = block;
    out->bytes_per_site_step_x8 = 2u * out->num_planes + (g->cfg.model != LGCA_B200_HPP ? 1u : 0u) + (out->has_no_slip ? 1u : 0u) +
                                  (out->has_slip ? 1u : 0u);
    return 0;
}

} // extern "C"
