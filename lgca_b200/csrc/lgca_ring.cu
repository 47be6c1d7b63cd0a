// Native halo ring for row strips: ghost rows are written straight into the ring neighbours' memory by a
// kernel (peer stores over NVLink / CUDA IPC mappings), ordered with device-side epoch flags -- no host
// synchronisation, no NCCL call and no packing buffers per block of fused steps.
//
// The reference has no domain decomposition; its torus is always periodic in y
// (src/omp_lattice.cpp:150-176), so strips form a periodic ring: upper neighbour = next strip in +y.
//
// Protocol per block b of <= k_fuse steps (epoch numbers count completed exchanges):
//     wait   : spin until both neighbours have published epoch b      (their ghost rows for this block are in)
//     step   : fused-step kernel reads buffer A (incl. ghost rows), writes the OWNED rows of buffer B
//     push   : copy my top/bottom `halo` owned rows of B into the neighbours' ghost rows of THEIR buffer B
//     signal : publish epoch b+1 in both neighbours' flag words (system-scope release)
// Safety: a neighbour only starts reading its B ghost rows after my signal b+1; and I only overwrite its A ghost
// rows at the end of block b+1, which I started after its signal b+1, i.e. after its block-b kernel (the last
// reader of A) had finished.  The step kernel never stores ghost rows, so pushes cannot be clobbered.
//
// Three plane sets rotate per strip (zero-copy snapshot, lgca_internal.h).  Every strip of a lattice issues the same
// sequence of steps / snapshots / in-place writes, so the live set has the same index on every strip and a push
// addresses the neighbour's set by that index.  With three sets the buffer a push overwrites was last read by a step
// kernel even earlier than in the two-buffer argument above; post-processing never reads ghost rows of the big
// buffers (the snapshot keeps a side copy of the one row it needs).
//
// Publishing WITHOUT a step in between (lgca_b200_ring_start / lgca_b200_ring_republish: after an upload, an
// initialiser or a body force changed the edge rows): the ghost rows being replaced may still be read by the
// neighbour's stream (its last step kernel, a snapshot side copy), so the publish first handshakes -- every strip
// posts an ACK ("my compute stream has passed every reader of epoch e") to both neighbours and waits for theirs.
#include <string.h>
#include <unistd.h>

#include <algorithm>

#include "lgca_internal.h"

namespace lgca_b200 {

struct RingBlob { // what a rank publishes to its neighbours
    uint64_t           magic;
    int32_t            pid, device;
    uint32_t           rows, pitch, halo, nd;
    uint64_t           plane_stride;
    void*              raw_planes[3]; // valid inside the publishing process
    void*              raw_flags;
    cudaIpcMemHandle_t ipc_planes[3];
    cudaIpcMemHandle_t ipc_flags;
};
static const uint64_t RING_MAGIC = 0x4C47434152494E47ull; // "LGCARING"
static const int      RING_CHAIN_DEPTH = 2;                 // chained launches in a row on a strip (lgca_b200::ring_step_blocks)

// copies `halo` rows of every plane: src rows [src_row, src_row+halo) -> dst rows [dst_row, ...)
__global__ void __launch_bounds__(256) ring_push_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst_upper,
                                                        uint32_t* __restrict__ dst_lower, Geom g, int nd,
                                                        uint64_t upper_stride, uint64_t lower_stride, uint32_t lower_rows)
{
    // x = word in [0, halo*pitch) as uint4, y = plane, z = 0: top rows -> upper neighbour, 1: bottom rows -> lower
    const uint32_t n4 = g.halo * g.pitch / 4;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const int d = blockIdx.y;
    const size_t plane = (size_t)d * g.plane_stride;
    if (blockIdx.z == 0) {
        // my highest owned rows become the upper neighbour's lower ghost rows [0, halo)
        const uint4 v = reinterpret_cast<const uint4*>(src + plane + (size_t)(g.rows - 2 * g.halo) * g.pitch)[i];
        reinterpret_cast<uint4*>(dst_upper + (size_t)d * upper_stride)[i] = v;
    } else {
        // my lowest owned rows become the lower neighbour's upper ghost rows [rows-halo, rows)
        const uint4 v = reinterpret_cast<const uint4*>(src + plane + (size_t)g.halo * g.pitch)[i];
        reinterpret_cast<uint4*>(dst_lower + (size_t)d * lower_stride + (size_t)(lower_rows - g.halo) * g.pitch)[i] = v;
    }
}

__global__ void ring_signal_kernel(volatile uint32_t* upper_flag_from_lower, volatile uint32_t* lower_flag_from_upper,
                                   uint32_t epoch, volatile uint32_t* my_push_done)
{
    __threadfence_system();
    // the upper neighbour sees me as its LOWER neighbour (flag slot 0), the lower one as its UPPER (slot 1)
    *upper_flag_from_lower = epoch;
    *lower_flag_from_upper = epoch;
    // my own slot 4: the push of this epoch has read my edge rows (edge tiles of chained launches wait for it before
    // they overwrite those rows)
    if (my_push_done) *my_push_done = epoch;
    __threadfence_system();
}

__global__ void ring_wait_kernel(volatile uint32_t* my_flags, uint32_t epoch)
{
    // my_flags[0]: published by my lower neighbour, my_flags[1]: by my upper neighbour ([2], [3]: their acks)
    while (my_flags[threadIdx.x] < epoch) __nanosleep(200);
    __threadfence_system();
}

// push my edge rows of the live buffer to the neighbours and publish the next epoch, on stream `s`
static int ring_push_and_signal(lgca_b200_lattice* h, cudaStream_t s)
{
    const Geom& g = h->g;
    const int b = buffer_id(h, h->planes[h->cur]); // same index on every strip (lockstep rotation)
    if (b < 0) return set_error(LGCA_B200_ESTATE, "live plane set is not one of the handle's three buffers");
    dim3 grid((g.halo * g.pitch / 4 + 255) / 256, h->nd, 2);
    ring_push_kernel<<<grid, 256, 0, s>>>(h->planes[h->cur], (uint32_t*)h->ring_upper_planes[b], (uint32_t*)h->ring_lower_planes[b],
                                          g, h->nd, h->ring_upper_stride, h->ring_lower_stride, h->ring_lower_rows);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    h->ring_epoch++;
    ring_signal_kernel<<<1, 1, 0, s>>>((volatile uint32_t*)h->ring_upper_flags + 0, (volatile uint32_t*)h->ring_lower_flags + 1,
                                       h->ring_epoch, (volatile uint32_t*)h->ring_flags + 4);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ring_wait_current_epoch(lgca_b200_lattice* h)
{
    ring_wait_kernel<<<1, 2, 0, h->s_compute>>>((volatile uint32_t*)h->ring_flags, h->ring_epoch);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// the compute stream waits for my most recent ghost-row push (it reads the edge rows an in-place writer is about to change)
int ring_order_inplace_write(lgca_b200_lattice* h)
{
    if (!h->ring_connected || h->ring_blocks == 0) return 0;
    LGCA_CUDA_CHECK(cudaStreamWaitEvent(h->s_compute, h->ev_push[(h->ring_blocks - 1) & 1u], 0));
    return 0;
}

static int open_peer(const RingBlob& blob, int my_device, void* planes[3], void** flags)
{
    if (blob.magic != RING_MAGIC) return set_error(LGCA_B200_EINVAL, "not a ring descriptor");
    if (blob.pid == (int32_t)getpid()) {
        // same process (several strips on one box driven by one host process, tests): plain pointers
        if (blob.device != my_device) {
            int can = 0;
            LGCA_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, my_device, blob.device));
            if (!can) return set_error(LGCA_B200_ENODEV, "no peer access between devices %d and %d", my_device, blob.device);
            cudaError_t e = cudaDeviceEnablePeerAccess(blob.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_cuda_error(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
        for (int i = 0; i < 3; ++i) planes[i] = blob.raw_planes[i];
        *flags = blob.raw_flags;
        return 0;
    }
    for (int i = 0; i < 3; ++i) LGCA_CUDA_CHECK(cudaIpcOpenMemHandle(&planes[i], blob.ipc_planes[i], cudaIpcMemLazyEnablePeerAccess));
    LGCA_CUDA_CHECK(cudaIpcOpenMemHandle(flags, blob.ipc_flags, cudaIpcMemLazyEnablePeerAccess));
    return 1; // opened through IPC: must be closed
}

} // namespace lgca_b200

using namespace lgca_b200;

extern "C" {

int lgca_b200_ring_descriptor_bytes(size_t* bytes)
{
    if (!bytes) return set_error(LGCA_B200_EINVAL, "null argument");
    *bytes = sizeof(RingBlob);
    return 0;
}

int lgca_b200_ring_export(lgca_b200_lattice* h, void* descriptor, size_t bytes)
{
    if (!h || !descriptor) return set_error(LGCA_B200_EINVAL, "null argument");
    if (bytes < sizeof(RingBlob)) return set_error(LGCA_B200_EINVAL, "descriptor buffer too small");
    if (!h->g.halo) return set_error(LGCA_B200_ESTATE, "handle owns the whole lattice: no ring");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    if (!h->ring_flags) {
        LGCA_CUDA_CHECK(cudaMalloc(&h->ring_flags, 64));
        LGCA_CUDA_CHECK(cudaMemset(h->ring_flags, 0, 64));
        LGCA_CUDA_CHECK(cudaDeviceSynchronize()); // default-stream memset vs. the non-blocking ring streams
        // highest priority: the tiny push/signal kernels must be dispatched ahead of the next step kernel's blocks,
        // which become ready at the same moment and would otherwise fill every SM first
        int prio_lo = 0, prio_hi = 0;
        LGCA_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        LGCA_CUDA_CHECK(cudaStreamCreateWithPriority(&h->s_ring, cudaStreamNonBlocking, prio_hi));
        for (int i = 0; i < 2; ++i) {
            LGCA_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_step[i], cudaEventDisableTiming));
            LGCA_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_push[i], cudaEventDisableTiming));
        }
    }
    RingBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = RING_MAGIC;
    b.pid = (int32_t)getpid();
    b.device = h->cfg.device;
    b.rows = h->g.rows; b.pitch = h->g.pitch; b.halo = h->g.halo; b.nd = (uint32_t)h->nd;
    b.plane_stride = h->g.plane_stride;
    for (int i = 0; i < 3; ++i) {
        b.raw_planes[i] = h->base[i];
        LGCA_CUDA_CHECK(cudaIpcGetMemHandle(&b.ipc_planes[i], h->base[i]));
    }
    b.raw_flags = h->ring_flags;
    LGCA_CUDA_CHECK(cudaIpcGetMemHandle(&b.ipc_flags, h->ring_flags));
    memcpy(descriptor, &b, sizeof(b));
    return 0;
}

int lgca_b200_ring_connect(lgca_b200_lattice* h, const void* lower_descriptor, const void* upper_descriptor)
{
    if (!h || !lower_descriptor || !upper_descriptor) return set_error(LGCA_B200_EINVAL, "null argument");
    if (!h->g.halo || !h->ring_flags) return set_error(LGCA_B200_ESTATE, "call lgca_b200_ring_export on this handle first");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    RingBlob lo, up;
    memcpy(&lo, lower_descriptor, sizeof(lo));
    memcpy(&up, upper_descriptor, sizeof(up));
    for (const RingBlob* b : {&lo, &up}) {
        if (b->magic != RING_MAGIC) return set_error(LGCA_B200_EINVAL, "not a ring descriptor");
        if (b->pitch != h->g.pitch || b->halo != h->g.halo || b->nd != (uint32_t)h->nd)
            return set_error(LGCA_B200_EINVAL, "neighbour strip has a different geometry (pitch/halo/planes)");
    }
    // ghost-row destinations are addressed with the NEIGHBOUR's row count / plane stride (strips may differ in height)
    h->ring_lower_rows = lo.rows;
    h->ring_lower_stride = lo.plane_stride;
    h->ring_upper_stride = up.plane_stride;
    // every strip must rotate its buffers in lockstep: start from the same live index
    int rc = open_peer(lo, h->cfg.device, h->ring_lower_planes, &h->ring_lower_flags);
    if (rc < 0) return rc;
    h->ring_lower_ipc = rc;
    const bool same = memcmp(&lo, &up, sizeof(lo)) == 0; // world == 2: both neighbours are the same strip
    if (same) {
        for (int i = 0; i < 3; ++i) h->ring_upper_planes[i] = h->ring_lower_planes[i];
        h->ring_upper_flags = h->ring_lower_flags;
        h->ring_upper_ipc = 0;
    } else {
        rc = open_peer(up, h->cfg.device, h->ring_upper_planes, &h->ring_upper_flags);
        if (rc < 0) return rc;
        h->ring_upper_ipc = rc;
    }
    // every kernel the ring will enqueue must already be resident: a lazy module load behind a spinning wait
    // kernel of the same process can deadlock
    if ((rc = wave_prepare(h))) return rc;
    if ((rc = simple_prepare(h))) return rc;
    cudaFuncAttributes fa;
    LGCA_CUDA_CHECK(cudaFuncGetAttributes(&fa, ring_push_kernel));
    LGCA_CUDA_CHECK(cudaFuncGetAttributes(&fa, ring_signal_kernel));
    LGCA_CUDA_CHECK(cudaFuncGetAttributes(&fa, ring_wait_kernel));
    // my own push-done word restarts with the epochs (the neighbours' words in [0..3] are theirs to write)
    LGCA_CUDA_CHECK(cudaMemsetAsync((uint32_t*)h->ring_flags + 4, 0, sizeof(uint32_t), h->s_compute));
    LGCA_CUDA_CHECK(cudaStreamSynchronize(h->s_compute));
    h->ring_connected = 1;
    h->ring_epoch = 0;
    h->ring_chain_k = 0;
    h->ring_chain_run = 0;
    return 0;
}

// Publishes the current edge rows to the neighbours without a step in between: after upload / init (first call:
// epoch 1) and after anything that changed the live state in place (body force, a new upload).  Collective: every
// strip of the lattice must call it at the same point of its call sequence.  Stream-ordered, no host synchronisation:
//     wait for my last push -> ACK to both neighbours -> wait for their ACKs -> push -> signal (next epoch)
int lgca_b200_ring_republish(lgca_b200_lattice* h)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    if (!h->ring_connected) return set_error(LGCA_B200_ESTATE, "ring not connected");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    int rc = ring_order_inplace_write(h);
    if (rc) return rc;
    const uint32_t next = h->ring_epoch + 1;
    // my compute stream has passed every reader of my ghost rows (step kernels, snapshot side copies): tell the
    // neighbours they may overwrite them; slot 2 = ack from the lower neighbour, slot 3 = from the upper one
    ring_signal_kernel<<<1, 1, 0, h->s_compute>>>((volatile uint32_t*)h->ring_upper_flags + 2, (volatile uint32_t*)h->ring_lower_flags + 3, next, nullptr);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    ring_wait_kernel<<<1, 2, 0, h->s_compute>>>((volatile uint32_t*)h->ring_flags + 2, next);
    h->launches++;
    LGCA_CUDA_CHECK(cudaGetLastError());
    if ((rc = ring_push_and_signal(h, h->s_compute))) return rc;
    LGCA_CUDA_CHECK(cudaEventRecord(h->ev_push[h->ring_blocks & 1u], h->s_compute));
    h->ring_blocks++;
    return 0;
}

// First publish after upload / init and connect, before the first lgca_b200_ring_step (kept as its own entry point).
int lgca_b200_ring_start(lgca_b200_lattice* h) { return lgca_b200_ring_republish(h); }

} // extern "C"

// continue_chain: the caller guarantees that nothing was enqueued on this strip's compute stream since the last block of
// its previous ring_step_blocks call (lgca_group.cu: the block-major loop of one group_step)
int lgca_b200::ring_step_blocks(lgca_b200_lattice* h, int n_steps, bool continue_chain)
{
    if (!h) return set_error(LGCA_B200_EINVAL, "null handle");
    if (!h->ring_connected) return set_error(LGCA_B200_ESTATE, "ring not connected");
    if (h->ring_epoch == 0) return set_error(LGCA_B200_ESTATE, "call lgca_b200_ring_start first");
    if (n_steps < 0) return set_error(LGCA_B200_EINVAL, "n_steps < 0");
    LGCA_CUDA_CHECK(cudaSetDevice(h->cfg.device));
    // one kernel launch per exchange: the fused depth of the wavefront kernel, or 1 for the generic kernel
    const int block = std::min(steps_per_launch(h, h->k_fuse), (int)h->g.halo);
    int prev_k = continue_chain ? h->ring_chain_k : 0; // k of the in-kernel-wait wave launch enqueued right before (0 = none)
    while (n_steps > 0) {
        const int k = n_steps < block ? n_steps : block;
        const int slot = (int)(h->ring_blocks & 1u);
        const bool inkernel = !(h->cfg.flags & LGCA_B200_FLAG_SIMPLE_KERNEL) && wave_has_edge_chunks(h, k);
        // chained launch (lgca_step_wave.cu): the step kernel may start while the previous one drains, so nothing may sit
        // between the two on the compute stream but event records -- the WAR wait below moves into the edge tiles
        // At most RING_CHAIN_DEPTH launches in a row are chained, then one is stream-ordered again: tiles of chained
        // launches that (transitively) wait for a neighbour's push spin in warp slots, and the pushes themselves need a
        // slot on their GPU -- the bound keeps the spinners of a strip that runs ahead of its neighbours to a fraction
        // of the machine (bands x depth x (depth + 1) tiles), whatever the neighbours' host threads are doing.
        const bool chain = inkernel && prev_k == k && h->chain_done[k] != nullptr && h->ring_chain_run < RING_CHAIN_DEPTH;
        h->ring_chain_run = chain ? h->ring_chain_run + 1 : 0;
        // (WAR) this block overwrites edge rows that an earlier push read (pushes complete in order on s_ring)
        if (!chain && h->ring_blocks >= 2) LGCA_CUDA_CHECK(cudaStreamWaitEvent(h->s_compute, h->ev_push[slot], 0));
        int rc;
        if (inkernel) {
            // tiles that read ghost rows wait in-kernel; the rest of the strip starts immediately
            h->ring_inkernel_epoch = h->ring_epoch;
            rc = step_one_launch(h, k, chain);
            h->ring_inkernel_epoch = 0;
            prev_k = k;
        } else {
            prev_k = 0;
            ring_wait_kernel<<<1, 2, 0, h->s_compute>>>((volatile uint32_t*)h->ring_flags, h->ring_epoch);
            h->launches++;
            LGCA_CUDA_CHECK(cudaGetLastError());
            rc = lgca_b200_step(h, k);
        }
        if (rc) return rc;
        // push + signal overlap with the next block's interior tiles
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_step[slot], h->s_compute));
        LGCA_CUDA_CHECK(cudaStreamWaitEvent(h->s_ring, h->ev_step[slot], 0));
        if ((rc = ring_push_and_signal(h, h->s_ring))) return rc;
        LGCA_CUDA_CHECK(cudaEventRecord(h->ev_push[slot], h->s_ring));
        h->ring_blocks++;
        n_steps -= k;
    }
    h->ring_chain_k = prev_k;
    return 0;
}

extern "C" {

int lgca_b200_ring_step(lgca_b200_lattice* h, int n_steps) { return lgca_b200::ring_step_blocks(h, n_steps, false); }

int lgca_b200_ring_disconnect(lgca_b200_lattice* h)
{
    if (!h) return 0;
    if (!h->ring_connected) return 0;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->s_compute);
    if (h->s_ring) cudaStreamSynchronize(h->s_ring);
    if (h->ring_lower_ipc) {
        for (int i = 0; i < 3; ++i) cudaIpcCloseMemHandle(h->ring_lower_planes[i]);
        cudaIpcCloseMemHandle(h->ring_lower_flags);
    }
    if (h->ring_upper_ipc) {
        for (int i = 0; i < 3; ++i) cudaIpcCloseMemHandle(h->ring_upper_planes[i]);
        cudaIpcCloseMemHandle(h->ring_upper_flags);
    }
    cudaGetLastError();
    h->ring_connected = 0;
    return 0;
}

} // extern "C"
