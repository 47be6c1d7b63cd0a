// microbenchmark: LL-protocol ping-pong between CTA 0 and CTA b (different SMs): cycles per round trip
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void st_msg(uint2* p, uint32_t a, uint32_t b) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void st_weak(uint2* p, uint32_t a, uint32_t b) { asm volatile("st.global.cg.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ uint2 ld_msg(const uint2* p) { uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
// MODE 0: LL strong; 1: LL + __threadfence after the store; 2: weak store; 3: strong, with a __syncthreads before/after like the kernel; 
template <int MODE>
__global__ void k(uint2* buf, long long* out, int rounds, int partner, int nmsg_per_thread)
{
    const int tid = threadIdx.x, T = blockDim.x;
    if (blockIdx.x != 0 && blockIdx.x != partner) return;
    const bool me0 = blockIdx.x == 0;
    uint2* mine  = buf + (me0 ? 0 : 65536);    // I write here
    uint2* their = buf + (me0 ? 65536 : 0);    // I poll here
    long long t0 = clock64();
    for (int r = 1; r <= rounds; ++r) {
        if (me0) {
            for (int i = 0; i < nmsg_per_thread; ++i) { if (MODE == 2) st_weak(mine + tid + i * T, r, r); else st_msg(mine + tid + i * T, r, r); }
            if (MODE == 1) __threadfence();
            for (int i = 0; i < nmsg_per_thread; ++i) { uint2 v; do { v = ld_msg(their + tid + i * T); } while (v.y != (uint32_t)r); }
            if (MODE == 3) __syncthreads();
        } else {
            for (int i = 0; i < nmsg_per_thread; ++i) { uint2 v; do { v = ld_msg(their + tid + i * T); } while (v.y != (uint32_t)r); }
            if (MODE == 3) __syncthreads();
            for (int i = 0; i < nmsg_per_thread; ++i) { if (MODE == 2) st_weak(mine + tid + i * T, r, r); else st_msg(mine + tid + i * T, r, r); }
            if (MODE == 1) __threadfence();
        }
    }
    long long t1 = clock64();
    if (tid == 0 && me0) out[0] = (t1 - t0) / rounds;
}
int main()
{
    uint2* buf; long long* out; cudaMalloc(&buf, 2 * 65536 * 8 * 2); cudaMallocManaged(&out, 64);
    const char* names[] = {"LL strong", "LL strong + threadfence", "LL weak store (cg)", "LL strong + syncthreads"};
#define RUN(M, T, P, N) cudaMemset(buf, 0, 2 * 65536 * 8 * 2); k<M><<<148, T>>>(buf, out, 2000, P, N); cudaDeviceSynchronize(); printf("%-26s threads=%4d partner CTA %3d msgs/thread %d: %lld cycles per round trip  err=%d\n", names[M], T, P, N, out[0], (int)cudaGetLastError());
    for (int P : {1, 2, 74, 147}) { RUN(0, 32, P, 1) }
    RUN(0, 1024, 1, 1) RUN(0, 1024, 1, 4) RUN(1, 1024, 1, 1) RUN(2, 1024, 1, 1) RUN(3, 1024, 1, 1) RUN(3, 1024, 1, 4) RUN(3, 256, 1, 4) RUN(0, 32, 1, 8)
    return 0;
}
