// Internal (library-private) declarations: the handle behind lgca_b200_lattice and the kernel
// launchers implemented in the individual .cu files.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <cuda_runtime.h>

#include "../../include/lgca_b200.h"
#include "lgca_common.cuh"

#define LGCA_MAX_K 8      /* HPP: up to 8 fused steps; FHP: up to LGCA_MAX_K_FHP (register budget) */
#define LGCA_MAX_K_FHP 6

namespace lgca_b200 {
// tiling of the wavefront kernel for one k (lgca_step_wave.cu)
struct WavePlan {
    int bands;      // 30-word bands per row
    int chunk_rows; // output rows per chunk (even)
    int edge_rows;  // strips: rows of the bottom / top edge chunk (0 = uniform chunks)
    int chunks;     // chunks over the stored rows
    int tiles;      // bands * chunks = warps launched
};
} // namespace lgca_b200

struct lgca_b200_lattice {
    lgca_b200_config cfg;
    lgca_b200::Geom  g;
    int              nd;          // NUM_DIR of the model
    int              k_fuse;      // steps fused per pass by the wavefront kernel
    uint32_t*        planes[2];   // ping-pong occupation planes [nd][rows][pitch]
    int              cur;         // index of the live buffer
    uint32_t*        snap;        // snapshot planes (copy_data_to_output_buffer)
    // Whole lattices snapshot WITHOUT a copy: the live buffer itself becomes the snapshot (snap == planes[cur],
    // `snap_spare` = the retired snapshot buffer).  The next step still reads it and writes the other buffer; after
    // that step the spare takes the shared buffer's slot in the ping-pong pair.  Anything that writes the live buffer
    // in place first calls unalias_snapshot() (copy-on-write).  Strips keep the copy: the ring neighbours hold fixed
    // mappings of planes[0] / planes[1].
    uint32_t*        snap_spare;
    // The three plane sets behind planes[0] / planes[1] / snap in allocation order.  They rotate (zero-copy snapshot),
    // so ring neighbours address a peer's buffer by its index in this array: all strips of a lattice perform the same
    // sequence of steps / snapshots / in-place writes, hence the live buffer has the same index on every strip.
    uint32_t*        base[3];
    // strips: the one row beyond the owned rows that the coarse means of the top coarse row read (first row of the
    // upper neighbour), copied at snapshot time so that post-processing never reads ghost rows of a rotating buffer
    uint32_t*        snap_ghost;  // [nd][pitch]
    int              sm_count;    // multiProcessorCount of the handle's device
    uint32_t*        ns;          // no-slip solid mask plane  [rows][pitch]
    uint32_t*        sl;          // slip solid mask plane
    uint32_t*        ch;          // chirality plane
    uint32_t*        xedge;       // [pitch] mask of the E/W domain-edge sites of a row
    uint32_t*        d_flags;     // [2] device flags: any no-slip / any slip cell
    int              has_ns, has_sl;
    int              have_state, have_types, have_rnd;
    cudaStream_t     s_compute, s_post;
    cudaStream_t     s_copy;                 // PCIe copies of upload/download, pipelined against pack/unpack
    cudaEvent_t      ev_stage_free[2], ev_stage_full[2];
    cudaEvent_t      ev_snap, ev_post, ev_t0, ev_t1;
    // Two host threads may drive one handle (stepping + snapshot on one, post_process / mean_velocity on the other,
    // apps/pipe/pipe_viewer.cpp:105,150).  The device-side ordering between "snapshot overwrites the buffer" and
    // "the post stream reads it" rests on ev_snap / ev_post, and an event wait only sees records that were enqueued
    // BEFORE it: this mutex makes "wait + enqueue + record" atomic on the host for both sides.
    pthread_mutex_t  snap_mutex;
    int              snap_mutex_init;
    // staging (lazily allocated, reused)
    void*            d_stage[2];
    size_t           stage_bytes;
    float*           d_cell_density;
    float*           d_cell_momentum;
    float*           d_mean_density;
    float*           d_mean_momentum;
    double*          d_scalars;   // small reduction scratch (device)
    double*          h_scalars;   // pinned host mirror
    int32_t*         d_draws;     // body-force scratch
    uint8_t*         d_draw_bytes;
    uint8_t*         h_draw_bytes;
    size_t           draw_cap;
    lgca_b200::WavePlan plans[LGCA_MAX_K + 1];
    int              plan_valid[LGCA_MAX_K + 1];
    uint8_t*         tile_fluid[LGCA_MAX_K + 1];      // per plan: 1 = tile window holds no solid site (wall variants)
    size_t           tile_fluid_cap[LGCA_MAX_K + 1];
    uint32_t*        chain_done[LGCA_MAX_K + 1];      // per plan: per-chunk completion counters of chained launches (whole lattices)
    size_t           chain_cap[LGCA_MAX_K + 1];
    uint32_t         chain_launches[LGCA_MAX_K + 1];  // launches of the plan since its counters were zeroed
    uint64_t         launches;
    uint64_t         device_bytes;
    // native halo ring (lgca_ring.cu)
    void*            ring_flags;            // [0..1] epochs published by my lower / upper neighbour, [2..3] their acks
    void*            ring_lower_planes[3];  // the lower neighbour's three plane sets (peer / IPC mapping), by base[] index
    void*            ring_upper_planes[3];
    uint32_t         ring_lower_rows;       // stored rows / plane stride of the neighbours (strips may differ in height)
    uint64_t         ring_lower_stride, ring_upper_stride;
    void*            ring_lower_flags;
    void*            ring_upper_flags;
    int              ring_lower_ipc, ring_upper_ipc, ring_connected;
    uint32_t         ring_epoch;
    uint64_t         ring_blocks;           // blocks issued since ring_start
    uint32_t         ring_inkernel_epoch;   // != 0: the next wave launch waits in-kernel for this epoch
    int              ring_chain_k;          // k of the wave launch that ended the last ring_step_blocks call (0 = not chainable)
    int              ring_chain_run;        // chained launches in a row so far
    cudaStream_t     s_ring;                // pushes + signals run here, overlapped with the next step kernel
    cudaEvent_t      ev_step[2], ev_push[2];
    // SM-resident kernel (lgca_step_resident.cu): ghost-row exchange area between CTAs and their progress counters
    uint32_t*        res_exch;              // {word, tag} messages
    size_t           res_exch_words;
    uint32_t         res_epoch;             // message tags are monotonic across launches
    // device-side exact body force (lgca_bodyforce.cu): draws, first-occurrence hash table, gains, block sums
    int32_t*         d_bf_draws;
    uint32_t*        d_bf_keys;
    uint8_t*         d_bf_gain;
    uint32_t*        d_bf_blocks;
    size_t           bf_cap;
    uint8_t*         d_bf_peer;             // first strip of a group: another strip's gains (peer copy)
    size_t           bf_peer_cap;
    cudaEvent_t      ev_bf;                 // end of this strip's classification
    // order-exact mean velocity (lgca_mv.cu): rounding tables, per-segment summaries, class bytes + pinned mirrors
    int32_t*         d_mv_tab;
    int32_t*         d_mv_rec;
    uint32_t*        d_mv_fluid;
    uint8_t*         d_mv_cls;
    int32_t*         h_mv_rec;
    uint32_t*        h_mv_fluid;
    uint8_t*         h_mv_cls;
    uint64_t         mv_stats[4];           // segments taken as one integer add / walked cell by cell; ns on the device + copies / in the walk
};

namespace lgca_b200 {

struct SnapLock { // scoped lock of snap_mutex (early returns of the CUDA-check macros release it)
    pthread_mutex_t* m;
    explicit SnapLock(lgca_b200_lattice* h) : m(&h->snap_mutex) { pthread_mutex_lock(m); }
    ~SnapLock() { if (m) pthread_mutex_unlock(m); }
    void release() { if (m) { pthread_mutex_unlock(m); m = nullptr; } }
    SnapLock(const SnapLock&) = delete;
    SnapLock& operator=(const SnapLock&) = delete;
};

// lgca_step_simple.cu : one 32-site word per thread, one step per pass
int launch_step_simple(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, cudaStream_t s);
// lgca_step_wave.cu : register wavefront, k steps per HBM pass
// chain = the operation enqueued on `s` right before this one is a launch_step_wave of the same handle and k: the launch
// may start while that one drains (per-chunk completion counters order the data; lgca_step_wave.cu)
int launch_step_wave(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int k, cudaStream_t s, bool chain = false);
bool wave_supported(const lgca_b200_lattice* h, int k);
// lgca_step_resident.cu : the whole lattice in shared memory, n steps per launch (lattices of up to ~20 MB of planes)
bool resident_supported(const lgca_b200_lattice* h);
int launch_step_resident(lgca_b200_lattice* h, const uint32_t* in, uint32_t* out, int n_steps, cudaStream_t s);
int resident_info(const lgca_b200_lattice* h, int* ctas, int* steps_per_exchange, size_t* smem_bytes);
int wave_prepare(lgca_b200_lattice* h);
bool wave_has_edge_chunks(lgca_b200_lattice* h, int k);
int simple_prepare(lgca_b200_lattice* h);
int ring_step_blocks(lgca_b200_lattice* h, int n_steps, bool continue_chain); // lgca_ring.cu: lgca_b200_ring_step + chain hint
int step_one_launch(lgca_b200_lattice* h, int k, bool chain); // lgca_capi.cu: k <= steps_per_launch steps of a strip in ONE wave launch
int ring_wait_current_epoch(lgca_b200_lattice* h); // lgca_ring.cu: stream-ordered wait for the neighbours' latest pushes
int ring_order_inplace_write(lgca_b200_lattice* h); // lgca_ring.cu: the compute stream waits for my last ghost-row push
int unalias_snapshot(lgca_b200_lattice* h); // lgca_capi.cu: give the live state a buffer of its own before an in-place write
int mean_velocity_sums(lgca_b200_lattice* h, double out3[3]); // lgca_capi.cu
void free_mv_buffers(lgca_b200_lattice* h); // lgca_mv.cu
void free_bf_buffers(lgca_b200_lattice* h); // lgca_bodyforce.cu
int body_force_device(lgca_b200_lattice* h, uint32_t forcing, bool first, const int32_t* draws, size_t n, size_t* consumed,
                      uint32_t* reverted); // lgca_bodyforce.cu: one batch, whole-lattice handles
// ... and its four stages over row strips (lgca_group.cu drives them)
int body_force_classify(lgca_b200_lattice* h, const int32_t* draws, size_t n);
int body_force_combine(lgca_b200_lattice* h0, lgca_b200_lattice* other, size_t n);
int body_force_cutoff(lgca_b200_lattice* h0, uint32_t forcing, bool first, size_t n, size_t* consumed, uint32_t* reverted);
int body_force_apply_cut(lgca_b200_lattice* h, size_t n, size_t consumed);
int steps_per_launch(const lgca_b200_lattice* h, int want); // lgca_capi.cu: steps ONE kernel launch can advance (<= want)
inline int buffer_id(const lgca_b200_lattice* h, const uint32_t* p)
{
    return p == h->base[0] ? 0 : (p == h->base[1] ? 1 : (p == h->base[2] ? 2 : -1));
}

// lgca_pack.cu : reference layouts <-> bit-planes
int launch_pack_state(lgca_b200_lattice* h, const uint8_t* d_bytes, uint32_t* planes, uint32_t row0, uint32_t nrows,
                      cudaStream_t s);
int launch_unpack_state(lgca_b200_lattice* h, const uint32_t* planes, uint8_t* d_bytes, uint32_t row0, uint32_t nrows,
                        cudaStream_t s);
int launch_pack_cell_type(lgca_b200_lattice* h, const int32_t* d_ct, uint32_t row0, uint32_t nrows, cudaStream_t s);
int launch_pack_rnd(lgca_b200_lattice* h, const uint8_t* d_bits, uint64_t first_bit, uint32_t row0, uint32_t nrows,
                    cudaStream_t s);
int launch_build_xedge(lgca_b200_lattice* h, cudaStream_t s);

// lgca_post.cu : popcount reductions and field kernels
int launch_cell_fields(lgca_b200_lattice* h, const uint32_t* planes, float* d_rho, float* d_mom, cudaStream_t s);
int launch_mean_fields(lgca_b200_lattice* h, const uint32_t* planes, const uint32_t* ghost_row, float* d_mrho, float* d_mmom,
                       int exact, cudaStream_t s);
int launch_mean_velocity(lgca_b200_lattice* h, const uint32_t* planes, double* d_out3, cudaStream_t s);
int launch_count_particles(lgca_b200_lattice* h, const uint32_t* planes, unsigned long long* d_out, cudaStream_t s);
int launch_gather_cells(lgca_b200_lattice* h, const uint32_t* planes, const int32_t* d_cells, size_t n, uint8_t* d_bytes,
                        cudaStream_t s);
int launch_apply_flips(lgca_b200_lattice* h, uint32_t* planes, const int32_t* d_cells, size_t n, cudaStream_t s);

// lgca_init.cu : device-side synthetic initial data and BC painting
int launch_init_random(lgca_b200_lattice* h, uint32_t* planes, uint64_t seed, cudaStream_t s);
int launch_paint_bc(lgca_b200_lattice* h, int bc_kind, cudaStream_t s);

} // namespace lgca_b200
