/*
 * lgca_b200.h -- C-ABI of the B200-native LGCA engine (liblgca_b200.so).
 *
 * This is the drop-in boundary for the hot path of keva92/lgca: everything the reference's
 * `Lattice<Model>` backend interface (src/lattice.h:32-238; the five pure virtuals at :187-203 and the
 * three copy hooks at :206-211, implemented on the CPU by OMP_Lattice, src/omp_lattice.cpp) needs from
 * a device backend, as plain C: opaque handle, plain pointers and sizes, int status codes.  The C++
 * `B200_Lattice<Model>` (lgca_b200/host/b200_lattice.h) is the binding a reference maintainer would
 * add next to OMP_Lattice; INTEGRATION.md shows it.
 *
 * Host-side array layouts are the reference's own (so existing apps keep working unchanged):
 *   state      uint8[dim_x*dim_y]      bit d of byte `cell` = occupation of direction d
 *                                      (src/omp_lattice.cpp:179,237; src/lgca_bitset.h:229-231)
 *   cell_type  int32[dim_x*dim_y]      0 FLUID, 1 SOLID_NO_SLIP, 2 SOLID_SLIP (src/lgca_common.h:53-57)
 *   rnd_bits   uint8[ceil(cells/8)]    chirality bit of cell i = bit i%8 of byte i/8
 *                                      (src/lgca_bitset.h:220-231; read at src/omp_lattice.cpp:198)
 *   cell = y*dim_x + x; y = 0 is the southern row.
 * On the device the lattice lives as bit-planes (one 32-bit word = 32 sites of one direction); that
 * layout is private to the library.
 *
 * Every function returns 0 on success and a negative LGCA_B200_E* code on failure;
 * lgca_b200_last_error() returns a message for the calling thread.  There is NO CPU fallback: without
 * a CUDA device every compute entry point fails with LGCA_B200_ENODEV.
 *
 * Threading (reference: apps/pipe/pipe_viewer.cpp:105,150 runs stepping and post-processing on two
 * host threads): {step, body force, snapshot, upload, download} use the handle's compute stream,
 * {post_process, mean_velocity, count on snapshot} use its post-processing stream; the snapshot buffer
 * separates them and event ordering makes the pair safe to call concurrently from two threads.
 */
#ifndef LGCA_B200_H_
#define LGCA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LGCA_B200_VERSION 1

/* models: enum class Model, src/lgca_common.h:46-51 */
enum { LGCA_B200_HPP = 0, LGCA_B200_FHP_I = 1, LGCA_B200_FHP_II = 2, LGCA_B200_FHP_III = 3 };

/* status codes */
enum {
    LGCA_B200_OK      = 0,
    LGCA_B200_EINVAL  = -1, /* bad argument */
    LGCA_B200_ENODEV  = -2, /* no CUDA device / wrong architecture: the product has no CPU path */
    LGCA_B200_ECUDA   = -3, /* CUDA runtime error (message in last_error) */
    LGCA_B200_ENOMEM  = -4,
    LGCA_B200_ESTATE  = -5  /* call order violated (e.g. step before upload) */
};

/* lgca_b200_config.flags */
enum {
    LGCA_B200_FLAG_NO_CELL_FIELDS = 1u << 0, /* never produce per-cell float fields (>= 1e9-cell runs) */
    LGCA_B200_FLAG_SIMPLE_KERNEL  = 1u << 1, /* force the one-word-per-thread kernel (debug / A-B tests) */
    LGCA_B200_FLAG_NO_RESIDENT    = 1u << 2, /* never use the SM-resident kernel (lattice kept in shared memory for all steps
                                                of a call); A-B tests of the HBM-streaming wavefront kernel on small lattices */
    LGCA_B200_FLAG_FORCE_RESIDENT = 1u << 3, /* use the SM-resident kernel whenever the lattice fits on chip, also where the
                                                library's own choice would be the wavefront kernel (A-B tests) */
    LGCA_B200_FLAG_HOST_BODY_FORCE = 1u << 4,/* lgca_b200_body_force through gather -> host replay -> apply instead of the
                                                device-side prefix (A-B tests; row strips always take this route) */
    LGCA_B200_FLAG_RESIDENT_DYNAMIC = 1u << 5,/* SM-resident kernel: dynamic four-word groups instead of static word ownership
                                                (A-B tests; the mapping of lattices with > 8192 words per CTA anyway) */
    LGCA_B200_FLAG_NO_CHAIN       = 1u << 6, /* wavefront kernel: consecutive launches of one lgca_b200_step call strictly one
                                                after the other instead of chained (the next launch filling the warp slots the
                                                previous one frees, ordered by per-chunk completion counters); A-B tests */
    LGCA_B200_FLAG_FORCE_CHAIN    = 1u << 7  /* chain launches also where one launch does not fill the machine (the library's
                                                own choice there is the plain stream order); tests of the chain on small lattices */
};

typedef struct lgca_b200_lattice lgca_b200_lattice; /* opaque */

typedef struct {
    int32_t  model;    /* LGCA_B200_HPP .. LGCA_B200_FHP_III */
    uint32_t dim_x;    /* global lattice width  (Lattice::m_dim_x, src/lattice.h:43) */
    uint32_t dim_y;    /* global lattice height (Lattice::m_dim_y; even for FHP, src/lattice.cpp:141) */
    uint32_t cg_radius;/* coarse graining radius (src/lattice.h:51); 0 = no coarse fields */
    int32_t  bf_dir;   /* body force direction 'x', 'y' or 0 (Lattice::m_bf_dir, src/lattice.cpp:125-133) */
    int32_t  device;   /* CUDA device ordinal */
    int32_t  k_fuse;   /* time steps fused per HBM pass (temporal blocking); 0 = library default */
    uint32_t y_begin;  /* first global row of the strip this handle owns (multi-GPU row strips) */
    uint32_t y_rows;   /* rows in the strip; 0 = the whole lattice (single GPU) */
    uint32_t flags;
} lgca_b200_config;

/* ---- life cycle: OMP_Lattice ctor/dtor + allocate_memory, src/omp_lattice.cpp:74-98,457-489 ---- */
int  lgca_b200_create(const lgca_b200_config* cfg, lgca_b200_lattice** out);
int  lgca_b200_destroy(lgca_b200_lattice* h);
const char* lgca_b200_last_error(void);
int  lgca_b200_version(void);
/* number of usable CUDA devices (0 when there is none; never fails) */
int  lgca_b200_device_count(void);

/* pinned host memory for the reference-layout mirrors (OMP_Lattice::allocate_memory uses malloc,
 * src/omp_lattice.cpp:457-471; pinned pages let uploads/downloads run at full PCIe speed) */
int  lgca_b200_host_alloc(size_t bytes, void** out);
int  lgca_b200_host_free(void* p);

/* ---- Lattice::copy_data_to_device(), src/lattice.h:206 / src/lattice.cpp:422-427 (empty hook) ----
 * Packs the reference-layout host arrays of the handle's strip into device bit-planes.  Any pointer
 * may be NULL to keep what is already on the device (state NULL => keep particles, etc.).  Host
 * arrays cover the STRIP's rows only: state/cell_type index = (y - y_begin)*dim_x + x; rnd_bits is
 * always indexed by the GLOBAL cell number (it is one flat bit-field in the reference). */
int  lgca_b200_upload(lgca_b200_lattice* h, const uint8_t* state, const int32_t* cell_type,
                      const uint8_t* rnd_bits);
/* ---- Lattice::copy_data_from_device(), src/lattice.h:209 ---- unpack + download the strip's state */
int  lgca_b200_download(lgca_b200_lattice* h, uint8_t* state);

/* ---- Lattice::collide_and_propagate(), src/lattice.h:191 / src/omp_lattice.cpp:100-249 ----
 * n_steps successive updates (periodic pull-stream, then collide / bounce at the destination cell).
 * Asynchronous on the compute stream.  The reference's `p` argument is ignored there and absent here
 * (the chirality comes from the frozen rnd bit-field). */
int  lgca_b200_step(lgca_b200_lattice* h, int n_steps);

/* ---- Lattice::copy_data_to_output_buffer(), src/lattice.h:211 / src/lattice.cpp:437-441 ----
 * Freezes the live state for post-processing.  No copy is made: the handle rotates three plane sets (the live set
 * becomes the snapshot, the next step writes a fresh set; in-place writers copy on write).  Row strips rotate in
 * lockstep, so all strips of a lattice must issue the same sequence of step / snapshot / in-place-write calls. */
int  lgca_b200_snapshot(lgca_b200_lattice* h);

/* ---- Lattice::post_process(), src/lattice.h:203 / src/omp_lattice.cpp:349-454 ----
 * Computes from the SNAPSHOT: per-cell density / momentum (AoS x,y) and the coarse-grained means over
 * the reference's window.  NULL outputs are skipped.  `exact_order` != 0 reproduces the
 * reference's float32 summation order for mean momentum y (bit-exact); 0 uses the popcount reduction
 * (density and momentum x are bit-exact either way). Synchronous (results are in host memory on return). */
int  lgca_b200_post_process(lgca_b200_lattice* h, float* cell_density, float* cell_momentum,
                            float* mean_density, float* mean_momentum, int exact_order);

/* ---- Lattice::get_mean_velocity(), src/lattice.h:194 / src/omp_lattice.cpp:508-557 ----
 * Device reduction over the snapshot (double accumulation; NOT the reference's sequential float32
 * order -- the order-exact variant runs in B200_Lattice on the host fields, as the reference does). */
int  lgca_b200_mean_velocity(lgca_b200_lattice* h, float out[2]);
/* Order-exact variant: CONTINUES the reference's sequential float32 sums (sum_x_vel, sum_y_vel and the FLUID-cell
 * counter of src/omp_lattice.cpp:513-549, as they come out at ONE thread) over this handle's rows of the SNAPSHOT, in
 * cell order, from the running values in sums[] / *fluid_cells.  Start them at 0, call the strips of a lattice in y
 * order and divide at the end: mean_velocity[i] = sums[i] / (float)fluid_cells (:553-554).  The device reduces every
 * 1024-cell segment to integer summaries per float32 binade, the host walks the segments (see csrc/lgca_mv.cu);
 * bit-equal to the reference loop.  Synchronous; post-processing stream.  At most 2^28 cells per handle. */
int  lgca_b200_mean_velocity_exact(lgca_b200_lattice* h, float sums[2], uint64_t* fluid_cells);
/* diagnostic counters of the calls so far: segments taken as one integer add, segments walked cell by cell,
 * nanoseconds spent on the device + copies, nanoseconds spent in the host walk */
int  lgca_b200_mean_velocity_stats(lgca_b200_lattice* h, uint64_t out[4]);
/* Host-only (no GPU) restatement of the same algorithm on a row-major array of class bytes (state byte of a FLUID
 * cell, 0 for solid cells), with the segment summaries computed on the CPU: the checker of the walk logic in the CPU
 * test-suite.  segments_fast / segments_walked (optional) count how the segments were taken. */
int  lgca_b200_mean_velocity_replay(int model, const uint8_t* class_bytes, uint32_t dim_x, uint32_t rows, float sums[2],
                                    uint64_t* segments_fast, uint64_t* segments_walked);

/* ---- Lattice::apply_body_force(), src/lattice.h:200 / src/omp_lattice.cpp:254-346 ----
 * Exact, draw-order-preserving body force.  `draws` are the caller's `rand()` values in stream order
 * (each is reduced `% num_cells` like src/omp_lattice.cpp:269).  Draws are consumed in order until
 * `forcing` particles have been reverted or `n_draws` are used up; *consumed and *reverted report the
 * progress so the caller can continue with more draws (the reference stops after 2*num_cells draws;
 * that cap is the caller's, see B200_Lattice::apply_body_force).  Operates on the live state.
 * Whole-lattice handles run the batch entirely on the device (csrc/lgca_bodyforce.cu: first occurrence of every
 * drawn cell through a hash table, gains from the bit-planes, prefix sum, the do-while's stop rule, scatter); the
 * only host synchronisation is the read-back of *consumed / *reverted. */
int  lgca_b200_body_force(lgca_b200_lattice* h, int forcing, const int32_t* draws, size_t n_draws,
                          size_t* consumed, uint32_t* reverted);

/* The three stages of the body force as separate calls, for drivers that own several strips (multi-GPU): gather the
 * bytes of the drawn cells on every strip (bit 7 set = not an eligible FLUID cell of this strip), combine them
 * (element-wise minimum over strips), replay the batch in draw order on the host (pure host function, no GPU), apply
 * the changed cells on every strip (cells outside a strip are ignored there). */
int  lgca_b200_body_force_gather(lgca_b200_lattice* h, const int32_t* cells, size_t n, uint8_t* bytes_out);
int  lgca_b200_body_force_replay(int model, int bf_dir, int forcing, const int32_t* cells, const uint8_t* bytes, size_t n,
                                 size_t* consumed, uint32_t* reverted, int32_t* changed_cells, uint8_t* changed_bytes,
                                 size_t* n_changed);
int  lgca_b200_body_force_apply(lgca_b200_lattice* h, const int32_t* cells, const uint8_t* new_bytes, size_t n);

/* ---- Lattice::get_n_particles(), src/lattice.h:165 / src/lattice.cpp:180-195 ---- (live state) */
int  lgca_b200_count_particles(lgca_b200_lattice* h, uint64_t* out);

/* ---- synthetic initial data for >= 1e8-cell throughput runs: occupancy with P = 1/NUM_DIR
 * in FLUID cells and chirality with P = 1/2 from a counter-based hash of (seed, cell, dir), generated
 * on the device.  Cell types must have been uploaded (or all-fluid via lgca_b200_fill_cell_type). ---- */
int  lgca_b200_init_random_device(lgca_b200_lattice* h, uint64_t seed);
/* paint a BC on the device for lattices too large to paint on the host:
 * bc = "periodic" | "pipe" | "karman" | "reflecting_back" | "reflecting_forward" (src/lattice.cpp:221-334) */
int  lgca_b200_apply_bc_device(lgca_b200_lattice* h, const char* bc);

/* ---- streams / timing ---- */
int  lgca_b200_sync(lgca_b200_lattice* h);
/* cudaStream_t of the compute stream, for callers that time with their own CUDA events */
void* lgca_b200_compute_stream(lgca_b200_lattice* h);
/* Runs n_steps updates bracketed by CUDA events on the compute stream; returns the elapsed device time. */
int  lgca_b200_timed_steps(lgca_b200_lattice* h, int n_steps, float* elapsed_ms);
/* Diagnostic for the roofline: `launches` back-to-back launches of the fused-step kernel (k_fuse steps each)
 * bracketed by CUDA events on the compute stream; returns the average launch duration.  On a row strip this
 * deliberately skips the halo exchange, so the strip's edge rows are stale afterwards (timing only). */
int  lgca_b200_timed_kernel(lgca_b200_lattice* h, int launches, float* ms_per_launch);
/* Kernel launches issued by this handle so far (bench.py's gpu_launches claim). */
int  lgca_b200_launch_count(lgca_b200_lattice* h, uint64_t* out);

/* ---- introspection ---- */
typedef struct {
    uint32_t dim_x, dim_y, y_begin, y_rows;
    uint32_t words_per_row;     /* 32-bit words per bit-plane row */
    uint32_t num_planes;        /* NUM_DIR */
    uint32_t has_no_slip, has_slip;
    int32_t  k_fuse;
    uint64_t bytes_per_site_step_x8; /* algorithmic bits per site per step: 2*NUM_DIR + mask planes */
    uint64_t device_bytes;      /* device memory held by the handle */
} lgca_b200_info;
int  lgca_b200_get_info(lgca_b200_lattice* h, lgca_b200_info* out);

/* ---- multi-GPU row strips (one handle per GPU / process) ----
 * The reference has no domain decomposition (single address space); this is the
 * B200-native extension.  Each strip keeps `halo` ghost rows above and below its own rows (halo = k_fuse
 * rounded up to even).  After every block of <= halo steps the owner exports its top and bottom `halo`
 * own rows and imports its ring neighbours' (periodic in y, like the reference's always-periodic torus,
 * src/omp_lattice.cpp:150-176).  Buffers are DEVICE pointers to packed rows ([plane][halo row][pitch
 * words]); the caller moves them between GPUs (NCCL send/recv, CUDA P2P).  Both calls are asynchronous
 * on the compute stream (lgca_b200_compute_stream) so the exchange can be stream-ordered without host
 * synchronisation.  `what` selects the occupation planes or the three static mask planes (no-slip,
 * slip, chirality), which are exchanged once after an upload from the host. */
enum { LGCA_B200_HALO_STATE = 0, LGCA_B200_HALO_MASKS = 1 };
int  lgca_b200_halo_rows(lgca_b200_lattice* h, uint32_t* rows);
/* Steps a strip may advance between two halo exchanges = steps ONE kernel launch can take: the fused depth of the
 * wavefront kernel, or 1 where only the generic kernel applies (dim_x < 64, fewer than 8 stored rows,
 * LGCA_B200_FLAG_SIMPLE_KERNEL).  lgca_b200_step rejects more than this on a strip (LGCA_B200_ESTATE). */
int  lgca_b200_steps_per_exchange(lgca_b200_lattice* h, int* steps);
int  lgca_b200_halo_bytes(lgca_b200_lattice* h, int what, size_t* bytes_per_side);
int  lgca_b200_halo_export(lgca_b200_lattice* h, int what, void* dev_top_rows, void* dev_bottom_rows);
int  lgca_b200_halo_import(lgca_b200_lattice* h, int what, const void* dev_from_upper, const void* dev_from_lower);
/* Wall-kind flags select the kernel variant.  With row strips every rank must run with the UNION of all
 * strips' flags (a strip without walls of its own may import wall cells into its halo rows). */
int  lgca_b200_get_wall_flags(lgca_b200_lattice* h, uint32_t* has_no_slip, uint32_t* has_slip);
int  lgca_b200_set_wall_flags(lgca_b200_lattice* h, uint32_t has_no_slip, uint32_t has_slip);

/* ---- native halo ring: ghost rows stored straight into the neighbours' memory (peer stores / CUDA IPC) ----
 * Each rank exports an opaque descriptor of its strip (lgca_b200_ring_descriptor_bytes bytes), the caller moves
 * the descriptors between processes (e.g. torch.distributed.all_gather), every rank connects to its lower
 * ((r-1) mod n) and upper ((r+1) mod n) neighbour, calls ring_start once after its data is in place, and
 * then advances with ring_step: per block of <= k_fuse steps the library enqueues wait -> fused-step kernel ->
 * peer push of the edge rows -> epoch signal on the compute stream; no host synchronisation, no collective.
 * All ranks must issue the same sequence of ring_step calls.  Strips must have equal heights. */
int  lgca_b200_ring_descriptor_bytes(size_t* bytes);
int  lgca_b200_ring_export(lgca_b200_lattice* h, void* descriptor, size_t bytes);
int  lgca_b200_ring_connect(lgca_b200_lattice* h, const void* lower_descriptor, const void* upper_descriptor);
int  lgca_b200_ring_start(lgca_b200_lattice* h);
int  lgca_b200_ring_step(lgca_b200_lattice* h, int n_steps);
/* Publishes the edge rows again WITHOUT a step in between: mandatory (on every strip, at the same point of the
 * call sequence) after anything that changed the live state in place while the ring is running -- the body force
 * (lgca_b200_body_force_apply), a new upload, an initialiser.  Stream-ordered; neighbours acknowledge that they no
 * longer read the ghost rows being replaced before the new rows are stored. */
int  lgca_b200_ring_republish(lgca_b200_lattice* h);
int  lgca_b200_ring_disconnect(lgca_b200_lattice* h);

/* ---- one lattice on several GPUs of one box, driven by ONE host process ----------------------------------------
 * The handle the C++ backend B200_Lattice<Model> talks to (reference boundary: the Lattice<Model> virtuals,
 * src/lattice.h:187-211; an app picks the backend where it says `new OMP_Lattice<MODEL>(...)`,
 * apps/pipe/pipe_viewer.cpp:46, or with the stale CLI's -p switch, apps/periodic/main.cpp:78-96).  A group owns
 * n_gpus row strips (contiguous, heights multiples of 2*cg_radius, even), one per device, connected by the native
 * halo ring through same-process peer pointers.  All host arrays are GLOBAL reference-layout arrays of the whole
 * lattice (see the top of this file); results are identical to a single-GPU handle (decomposition invariance).
 * n_gpus == 1 is a plain whole-lattice handle.  `dev_ids` NULL = devices 0 .. n_gpus-1; cfg->device, y_begin and
 * y_rows are ignored.  Same status codes and threading rules as the single-handle calls. */
typedef struct lgca_b200_group lgca_b200_group; /* opaque */
int  lgca_b200_group_create(const lgca_b200_config* cfg, int n_gpus, const int* dev_ids, lgca_b200_group** out);
int  lgca_b200_group_destroy(lgca_b200_group* g);
int  lgca_b200_group_size(lgca_b200_group* g, int* n_gpus);
/* borrowed pointer to strip i's handle (introspection, tests) */
int  lgca_b200_group_strip(lgca_b200_group* g, int i, lgca_b200_lattice** h);
int  lgca_b200_group_upload(lgca_b200_group* g, const uint8_t* state, const int32_t* cell_type, const uint8_t* rnd_bits);
int  lgca_b200_group_download(lgca_b200_group* g, uint8_t* state);
int  lgca_b200_group_step(lgca_b200_group* g, int n_steps);
int  lgca_b200_group_snapshot(lgca_b200_group* g);
int  lgca_b200_group_post_process(lgca_b200_group* g, float* cell_density, float* cell_momentum, float* mean_density,
                                  float* mean_momentum, int exact_order);
int  lgca_b200_group_mean_velocity(lgca_b200_group* g, float out[2]);
/* order-exact (lgca_b200_mean_velocity_exact chained over the strips in y order, then the final divisions) */
int  lgca_b200_group_mean_velocity_exact(lgca_b200_group* g, float out[2]);
int  lgca_b200_group_body_force(lgca_b200_group* g, int forcing, const int32_t* draws, size_t n_draws, size_t* consumed,
                                uint32_t* reverted);
int  lgca_b200_group_count_particles(lgca_b200_group* g, uint64_t* out);
int  lgca_b200_group_init_random_device(lgca_b200_group* g, uint64_t seed);
int  lgca_b200_group_apply_bc_device(lgca_b200_group* g, const char* bc);
int  lgca_b200_group_sync(lgca_b200_group* g);
/* n_steps updates bracketed by CUDA events on every strip's compute stream; elapsed = max over the strips */
int  lgca_b200_group_timed_steps(lgca_b200_group* g, int n_steps, float* elapsed_ms);
int  lgca_b200_group_launch_count(lgca_b200_group* g, uint64_t* out);
/* global dims; y_rows = dim_y; device_bytes summed over the strips; k_fuse = steps per halo exchange */
int  lgca_b200_group_get_info(lgca_b200_group* g, lgca_b200_info* out);

#ifdef __cplusplus
}
#endif
#endif /* LGCA_B200_H_ */
