// Exact, draw-order-preserving body force ON THE DEVICE (OMP_Lattice<M>::apply_body_force,
// src/omp_lattice.cpp:254-346).
//
// The reference is a sequential loop: draw cell = rand() % num_cells, revert one particle (two for FHP with bf 'y') if
// the cell is FLUID and eligible, stop once `forcing` particles are reverted.  A reverted cell is never eligible again
// (its target direction is now occupied), so the loop is equivalent to:
//     gain[i] = (draw i is the FIRST occurrence of its cell in the batch) ? #particles the untouched cell can revert : 0
//     P[i]    = gain[0] + ... + gain[i]
//     draw i is processed  <=>  i == 0 (do-while)  or  P[i-1] < forcing        (`unsigned < int` compares unsigned, :346)
// which is data-parallel: first occurrences through an open-addressing hash table (atomicCAS on the key, atomicMin on
// the draw index), gains from the bit-planes, a prefix sum over the batch, and a scatter of the processed reverts
// (atomicOr / atomicAnd on the plane words).  One stream-ordered sequence of four kernels per batch; the only host
// synchronisation is the read-back of { consumed, reverted }, which the caller's rand() FIFO needs.
#include "lgca_internal.h"

namespace lgca_b200 {

static constexpr uint32_t BF_EMPTY = 0xFFFFFFFFu;
static constexpr int      BF_BLOCK = 256;

__device__ __forceinline__ uint32_t bf_hash(uint32_t c) { return c * 2654435761u; }

__global__ void __launch_bounds__(BF_BLOCK) bf_insert_kernel(const int32_t* __restrict__ draws, uint32_t n, uint32_t num_cells,
                                                             uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t mask)
{
    const uint32_t i = blockIdx.x * BF_BLOCK + threadIdx.x;
    if (i >= n) return;
    const uint32_t cell = (uint32_t)draws[i] % num_cells; // :269
    uint32_t slot = (bf_hash(cell) >> 7) & mask;
    while (true) {
        const uint32_t prev = atomicCAS(keys + slot, BF_EMPTY, cell);
        if (prev == BF_EMPTY || prev == cell) { atomicMin(vals + slot, i); return; }
        slot = (slot + 1) & mask;
    }
}

// gain of an untouched cell with state byte b (src/omp_lattice.cpp:295-338)
template <int ND>
__device__ __forceinline__ uint32_t bf_gain(uint32_t b, int bf)
{
    if (ND == 4) {
        if (bf == 'x') return (!(b & 1u) && (b & 4u)) ? 1u : 0u;
        if (bf == 'y') return ((b & 2u) && !(b & 8u)) ? 1u : 0u;
        return 0u;
    }
    if (bf == 'x') return (!(b & 1u) && (b & 8u)) ? 1u : 0u;
    if (bf == 'y') return (((b & 2u) && !(b & 32u)) ? 1u : 0u) + (((b & 4u) && !(b & 16u)) ? 1u : 0u);
    return 0u;
}

// gains + their per-block sums
template <int ND>
__global__ void __launch_bounds__(BF_BLOCK) bf_classify_kernel(const uint32_t* __restrict__ planes, const uint32_t* __restrict__ ns,
                                                               const uint32_t* __restrict__ sl, const int32_t* __restrict__ draws,
                                                               uint32_t n, uint32_t num_cells, const uint32_t* __restrict__ keys,
                                                               const uint32_t* __restrict__ vals, uint32_t mask,
                                                               uint8_t* __restrict__ gain, uint32_t* __restrict__ block_sums,
                                                               const Geom g, uint32_t own_rows, int bf)
{
    const uint32_t i = blockIdx.x * BF_BLOCK + threadIdx.x;
    uint32_t gn = 0;
    if (i < n) {
        const uint32_t cell = (uint32_t)draws[i] % num_cells;
        uint32_t slot = (bf_hash(cell) >> 7) & mask;
        while (keys[slot] != cell) slot = (slot + 1) & mask;
        const uint32_t gy = cell / g.dim_x, x = cell % g.dim_x;
        if (vals[slot] == i && gy >= g.y0 && gy < g.y0 + own_rows) {
            const size_t   base = (size_t)(gy - g.y0 + g.halo) * g.pitch + (x >> 5);
            const uint32_t bit  = x & 31;
            if (!(((__ldg(ns + base) | __ldg(sl + base)) >> bit) & 1u)) {
                uint32_t b = 0;
#pragma unroll
                for (int d = 0; d < ND; ++d) b |= ((planes[(size_t)d * g.plane_stride + base] >> bit) & 1u) << d;
                gn = bf_gain<ND>(b, bf);
            }
        }
        gain[i] = (uint8_t)gn;
    }
    __shared__ uint32_t sh[BF_BLOCK / 32];
    uint32_t s = gn;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < BF_BLOCK / 32; ++w) t += sh[w];
        block_sums[blockIdx.x] = t;
    }
}

// exclusive scan of the block sums (one block; nblocks <= a few thousand)
__global__ void __launch_bounds__(1024) bf_scan_blocks_kernel(uint32_t* __restrict__ block_sums, uint32_t nblocks)
{
    __shared__ uint32_t sh[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? block_sums[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = sh[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o); if (threadIdx.x >= o) w += t; }
            sh[threadIdx.x] = w; // inclusive over warps
        }
        __syncthreads();
        const uint32_t incl = s + ((threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0u) + carry;
        if (i < nblocks) block_sums[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
}

// prefix inside the block, the stop rule, the scatter of the processed reverts, { consumed, reverted }
template <int ND>
__global__ void __launch_bounds__(BF_BLOCK) bf_apply_kernel(uint32_t* __restrict__ planes, const int32_t* __restrict__ draws,
                                                            uint32_t n, uint32_t num_cells, const uint8_t* __restrict__ gain,
                                                            const uint32_t* __restrict__ block_offsets, uint32_t forcing,
                                                            uint32_t first, uint32_t* __restrict__ out2, const Geom g, int bf)
{
    const uint32_t i = blockIdx.x * BF_BLOCK + threadIdx.x;
    const uint32_t gn = i < n ? gain[i] : 0u;
    __shared__ uint32_t sh[BF_BLOCK / 32];
    uint32_t s = gn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    uint32_t off = block_offsets[blockIdx.x];
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += sh[w];
    if (i >= n) return;
    const uint32_t incl = off + s, excl = incl - gn;
    const bool processed = (i == 0 && first) || excl < forcing;
    if (!processed) return;
    if (i == n - 1 || !(incl < forcing)) { out2[0] = i + 1; out2[1] = incl; } // the last processed draw
    if (!gn) return;
    const uint32_t cell = (uint32_t)draws[i] % num_cells;
    const uint32_t gy = cell / g.dim_x, x = cell % g.dim_x;
    const size_t   base = (size_t)(gy - g.y0 + g.halo) * g.pitch + (x >> 5);
    const uint32_t m = 1u << (x & 31);
    auto move = [&](int from, int to) {
        atomicAnd(planes + (size_t)from * g.plane_stride + base, ~m);
        atomicOr(planes + (size_t)to * g.plane_stride + base, m);
    };
    if (ND == 4) {
        if (bf == 'x') move(2, 0); else move(1, 3);
    } else if (bf == 'x') {
        move(3, 0);
    } else {
        const uint32_t b1 = (planes[(size_t)1 * g.plane_stride + base] & m) && !(planes[(size_t)5 * g.plane_stride + base] & m);
        const uint32_t b2 = (planes[(size_t)2 * g.plane_stride + base] & m) && !(planes[(size_t)4 * g.plane_stride + base] & m);
        if (b1) move(1, 5);
        if (b2) move(2, 4);
    }
}

static int ensure_bf_buffers(lgca_b200_lattice* h, size_t n)
{
    if (n <= h->bf_cap) return 0;
    cudaFree(h->d_bf_draws); cudaFree(h->d_bf_keys); cudaFree(h->d_bf_gain); cudaFree(h->d_bf_blocks);
    h->d_bf_draws = nullptr; h->d_bf_keys = nullptr; h->d_bf_gain = nullptr; h->d_bf_blocks = nullptr;
    h->bf_cap = 0;
    size_t cap = 1 << 16;
    while (cap < n) cap <<= 1;
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_bf_draws, cap * sizeof(int32_t)));
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_bf_keys, cap * 2 * 2 * sizeof(uint32_t))); // keys | vals, 2 slots per draw
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_bf_gain, cap));
    LGCA_CUDA_CHECK(cudaMalloc((void**)&h->d_bf_blocks, (cap / BF_BLOCK + 4) * sizeof(uint32_t)));
    h->bf_cap = cap;
    return 0;
}

void free_bf_buffers(lgca_b200_lattice* h)
{
    cudaFree(h->d_bf_draws); cudaFree(h->d_bf_keys); cudaFree(h->d_bf_gain); cudaFree(h->d_bf_blocks); cudaFree(h->d_bf_peer);
    if (h->ev_bf) cudaEventDestroy(h->ev_bf);
}

// One batch on a whole-lattice handle.  `first`: the batch opens a body-force call (the do-while's unconditional first
// draw).  Synchronous (returns consumed / reverted).
int body_force_device(lgca_b200_lattice* h, uint32_t forcing, bool first, const int32_t* draws, size_t n, size_t* consumed,
                      uint32_t* reverted)
{
    *consumed = 0;
    *reverted = 0;
    if (n == 0) return 0;
    if (n > 0x7FFFFFFFu) return set_error(LGCA_B200_EINVAL, "batch too large");
    const Geom& g = h->g;
    const uint32_t num_cells = (uint32_t)((uint64_t)g.dim_x * g.dim_y);
    int rc = ensure_bf_buffers(h, n);
    if (rc) return rc;
    cudaStream_t s = h->s_compute;
    size_t slots = 1 << 10;
    while (slots < 2 * n) slots <<= 1;
    uint32_t* keys = h->d_bf_keys;
    uint32_t* vals = h->d_bf_keys + slots;
    uint32_t* out2 = reinterpret_cast<uint32_t*>(h->d_scalars + 6);
    const uint32_t nb = (uint32_t)((n + BF_BLOCK - 1) / BF_BLOCK);
    const int bf = h->cfg.bf_dir;
    LGCA_CUDA_CHECK(cudaMemcpyAsync(h->d_bf_draws, draws, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    LGCA_CUDA_CHECK(cudaMemsetAsync(keys, 0xFF, 2 * slots * sizeof(uint32_t), s));
    LGCA_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(uint32_t), s)); // no draw processed (continuation batch, forcing met)
    if ((rc = unalias_snapshot(h))) return rc; // in-place write: the snapshot must not see it
    if ((rc = ring_order_inplace_write(h))) return rc;
    uint32_t* planes = h->planes[h->cur];
    const uint32_t own = g.rows - 2 * g.halo;
    bf_insert_kernel<<<nb, BF_BLOCK, 0, s>>>(h->d_bf_draws, (uint32_t)n, num_cells, keys, vals, (uint32_t)slots - 1);
#define BF_CLASSIFY(ND) bf_classify_kernel<ND><<<nb, BF_BLOCK, 0, s>>>(planes, h->ns, h->sl, h->d_bf_draws, (uint32_t)n, num_cells, keys, vals, \
                                                                       (uint32_t)slots - 1, h->d_bf_gain, h->d_bf_blocks, g, own, bf)
#define BF_APPLY(ND) bf_apply_kernel<ND><<<nb, BF_BLOCK, 0, s>>>(planes, h->d_bf_draws, (uint32_t)n, num_cells, h->d_bf_gain, h->d_bf_blocks, \
                                                                 forcing, first ? 1u : 0u, out2, g, bf)
    if (h->nd == 4) BF_CLASSIFY(4); else if (h->nd == 6) BF_CLASSIFY(6); else BF_CLASSIFY(7);
    bf_scan_blocks_kernel<<<1, 1024, 0, s>>>(h->d_bf_blocks, nb);
    if (h->nd == 4) BF_APPLY(4); else if (h->nd == 6) BF_APPLY(6); else BF_APPLY(7);
#undef BF_CLASSIFY
#undef BF_APPLY
    h->launches += 4;
    LGCA_CUDA_CHECK(cudaGetLastError());
    uint32_t* host2 = reinterpret_cast<uint32_t*>(h->h_scalars + 6);
    LGCA_CUDA_CHECK(cudaMemcpyAsync(host2, out2, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    *consumed = host2[0];
    *reverted = host2[1];
    return 0;
}

// ---- the same semantics over ROW STRIPS (lgca_group.cu) ---------------------------------------------------------------
// Every strip classifies the whole batch against its own rows (first occurrences are a property of the draw sequence, so
// every strip builds the same hash table; gains are non-zero on the owner only), the gains are summed on the first strip's
// device (peer copies), which runs the prefix sum and the stop rule; every strip then scatters its own reverts up to the cut.

__global__ void __launch_bounds__(BF_BLOCK) bf_combine_kernel(uint8_t* __restrict__ gain, const uint8_t* __restrict__ other, uint32_t n)
{
    const uint32_t i = blockIdx.x * BF_BLOCK + threadIdx.x;
    if (i < n) gain[i] = (uint8_t)(gain[i] | other[i]); // at most one strip owns the cell
}

__global__ void __launch_bounds__(BF_BLOCK) bf_block_sums_kernel(const uint8_t* __restrict__ gain, uint32_t n, uint32_t* __restrict__ block_sums)
{
    const uint32_t i = blockIdx.x * BF_BLOCK + threadIdx.x;
    uint32_t s = i < n ? gain[i] : 0u;
    __shared__ uint32_t sh[BF_BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < BF_BLOCK / 32; ++w) t += sh[w];
        block_sums[blockIdx.x] = t;
    }
}

// the stop rule alone: out2 = { consumed, reverted }
__global__ void __launch_bounds__(BF_BLOCK) bf_cutoff_kernel(const uint8_t* __restrict__ gain, uint32_t n, const uint32_t* __restrict__ block_offsets,
                                                             uint32_t forcing, uint32_t first, uint32_t* __restrict__ out2)
{
    const uint32_t i = blockIdx.x * BF_BLOCK + threadIdx.x;
    const uint32_t gn = i < n ? gain[i] : 0u;
    __shared__ uint32_t sh[BF_BLOCK / 32];
    uint32_t s = gn;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    uint32_t off = block_offsets[blockIdx.x];
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += sh[w];
    if (i >= n) return;
    const uint32_t incl = off + s, excl = incl - gn;
    if (!((i == 0 && first) || excl < forcing)) return;
    if (i == n - 1 || !(incl < forcing)) { out2[0] = i + 1; out2[1] = incl; }
}

// scatter of the strip's own reverts among the first `consumed` draws
template <int ND>
__global__ void __launch_bounds__(BF_BLOCK) bf_apply_cut_kernel(uint32_t* __restrict__ planes, const int32_t* __restrict__ draws, uint32_t consumed,
                                                                uint32_t num_cells, const uint8_t* __restrict__ gain, const Geom g,
                                                                uint32_t own_rows, int bf)
{
    const uint32_t i = blockIdx.x * BF_BLOCK + threadIdx.x;
    if (i >= consumed || !gain[i]) return;
    const uint32_t cell = (uint32_t)draws[i] % num_cells;
    const uint32_t gy = cell / g.dim_x, x = cell % g.dim_x;
    if (gy < g.y0 || gy >= g.y0 + own_rows) return; // another strip's cell (the first strip holds the combined gains)
    const size_t   base = (size_t)(gy - g.y0 + g.halo) * g.pitch + (x >> 5);
    const uint32_t m = 1u << (x & 31);
    auto move = [&](int from, int to) {
        atomicAnd(planes + (size_t)from * g.plane_stride + base, ~m);
        atomicOr(planes + (size_t)to * g.plane_stride + base, m);
    };
    if (ND == 4) {
        if (bf == 'x') move(2, 0); else move(1, 3);
    } else if (bf == 'x') {
        move(3, 0);
    } else {
        const uint32_t b1 = (planes[(size_t)1 * g.plane_stride + base] & m) && !(planes[(size_t)5 * g.plane_stride + base] & m);
        const uint32_t b2 = (planes[(size_t)2 * g.plane_stride + base] & m) && !(planes[(size_t)4 * g.plane_stride + base] & m);
        if (b1) move(1, 5);
        if (b2) move(2, 4);
    }
}

// stage 1 on one strip: draws -> device, hash table, gains of the strip's own cells (asynchronous; ev_bf marks the end)
int body_force_classify(lgca_b200_lattice* h, const int32_t* draws, size_t n)
{
    const Geom& g = h->g;
    const uint32_t num_cells = (uint32_t)((uint64_t)g.dim_x * g.dim_y);
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int rc = ensure_bf_buffers(h, n);
    if (rc) return rc;
    if (!h->ev_bf) LGCA_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_bf, cudaEventDisableTiming));
    cudaStream_t s = h->s_compute;
    size_t slots = 1 << 10;
    while (slots < 2 * n) slots <<= 1;
    uint32_t* keys = h->d_bf_keys;
    uint32_t* vals = h->d_bf_keys + slots;
    const uint32_t nb = (uint32_t)((n + BF_BLOCK - 1) / BF_BLOCK);
    const int bf = h->cfg.bf_dir;
    LGCA_CUDA_CHECK(cudaMemcpyAsync(h->d_bf_draws, draws, n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    LGCA_CUDA_CHECK(cudaMemsetAsync(keys, 0xFF, 2 * slots * sizeof(uint32_t), s));
    const uint32_t* planes = h->planes[h->cur];
    const uint32_t own = g.rows - 2 * g.halo;
    bf_insert_kernel<<<nb, BF_BLOCK, 0, s>>>(h->d_bf_draws, (uint32_t)n, num_cells, keys, vals, (uint32_t)slots - 1);
#define BF_CLASSIFY(ND) bf_classify_kernel<ND><<<nb, BF_BLOCK, 0, s>>>(planes, h->ns, h->sl, h->d_bf_draws, (uint32_t)n, num_cells, keys, vals, \
                                                                       (uint32_t)slots - 1, h->d_bf_gain, h->d_bf_blocks, g, own, bf)
    if (h->nd == 4) BF_CLASSIFY(4); else if (h->nd == 6) BF_CLASSIFY(6); else BF_CLASSIFY(7);
#undef BF_CLASSIFY
    h->launches += 2;
    LGCA_CUDA_CHECK(cudaGetLastError());
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_bf, s));
    return 0;
}

// stage 2 on the first strip: add another strip's gains (peer copy ordered behind that strip's classification)
int body_force_combine(lgca_b200_lattice* h0, lgca_b200_lattice* other, size_t n)
{
    LGCA_CUDA_CHECK(cudaSetDevice(h0->cfg.device));
    if (h0->bf_peer_cap < n) {
        cudaFree(h0->d_bf_peer);
        h0->d_bf_peer = nullptr; h0->bf_peer_cap = 0;
        LGCA_CUDA_CHECK(cudaMalloc((void**)&h0->d_bf_peer, h0->bf_cap));
        h0->bf_peer_cap = h0->bf_cap;
    }
    cudaStream_t s = h0->s_compute;
    LGCA_CUDA_CHECK(cudaStreamWaitEvent(s, other->ev_bf, 0));
    LGCA_CUDA_CHECK(cudaMemcpyPeerAsync(h0->d_bf_peer, h0->cfg.device, other->d_bf_gain, other->cfg.device, n, s));
    bf_combine_kernel<<<(unsigned)((n + BF_BLOCK - 1) / BF_BLOCK), BF_BLOCK, 0, s>>>(h0->d_bf_gain, h0->d_bf_peer, (uint32_t)n);
    h0->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// stage 3 on the first strip: prefix sum + stop rule over the combined gains (synchronous)
int body_force_cutoff(lgca_b200_lattice* h0, uint32_t forcing, bool first, size_t n, size_t* consumed, uint32_t* reverted)
{
    LGCA_CUDA_CHECK(cudaSetDevice(h0->cfg.device));
    cudaStream_t s = h0->s_compute;
    const uint32_t nb = (uint32_t)((n + BF_BLOCK - 1) / BF_BLOCK);
    uint32_t* out2 = reinterpret_cast<uint32_t*>(h0->d_scalars + 6);
    LGCA_CUDA_CHECK(cudaMemsetAsync(out2, 0, 2 * sizeof(uint32_t), s));
    bf_block_sums_kernel<<<nb, BF_BLOCK, 0, s>>>(h0->d_bf_gain, (uint32_t)n, h0->d_bf_blocks);
    bf_scan_blocks_kernel<<<1, 1024, 0, s>>>(h0->d_bf_blocks, nb);
    bf_cutoff_kernel<<<nb, BF_BLOCK, 0, s>>>(h0->d_bf_gain, (uint32_t)n, h0->d_bf_blocks, forcing, first ? 1u : 0u, out2);
    h0->launches += 3;
    LGCA_CUDA_CHECK(cudaGetLastError());
    uint32_t* host2 = reinterpret_cast<uint32_t*>(h0->h_scalars + 6);
    LGCA_CUDA_CHECK(cudaMemcpyAsync(host2, out2, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    LGCA_CUDA_CHECK(cudaStreamSynchronize(s));
    *consumed = host2[0];
    *reverted = host2[1];
    return 0;
}

// stage 4 on every strip: scatter (asynchronous on the strip's compute stream)
int body_force_apply_cut(lgca_b200_lattice* h, size_t n, size_t consumed)
{
    (void)n;
    if (consumed == 0) return 0;
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int rc;
    if ((rc = unalias_snapshot(h))) return rc; // in-place write: the snapshot must not see it
    if ((rc = ring_order_inplace_write(h))) return rc;
    const Geom& g = h->g;
    const uint32_t num_cells = (uint32_t)((uint64_t)g.dim_x * g.dim_y);
    const uint32_t own = g.rows - 2 * g.halo;
    const uint32_t nb = (uint32_t)((consumed + BF_BLOCK - 1) / BF_BLOCK);
    uint32_t* planes = h->planes[h->cur];
    const int bf = h->cfg.bf_dir;
    cudaStream_t s = h->s_compute;
#define BF_CUT(ND) bf_apply_cut_kernel<ND><<<nb, BF_BLOCK, 0, s>>>(planes, h->d_bf_draws, (uint32_t)consumed, num_cells, h->d_bf_gain, g, own, bf)
    if (h->nd == 4) BF_CUT(4); else if (h->nd == 6) BF_CUT(6); else BF_CUT(7);
#undef BF_CUT
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

} // namespace lgca_b200
