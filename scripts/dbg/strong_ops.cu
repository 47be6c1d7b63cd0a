// microbenchmark: issue cost of strong (relaxed.gpu) vs weak global stores / loads from one warp and from 32 warps
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void st_strong(uint2* p, uint32_t a, uint32_t b) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void st_strong4(uint4* p, uint32_t a, uint32_t b) { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void st_weak(uint2* p, uint32_t a, uint32_t b) { asm volatile("st.global.cg.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ uint2 ld_strong(const uint2* p) { uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint2 ld_weak(const uint2* p) { uint2 v; asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
template <int MODE, int N>
__global__ void k(uint2* buf, long long* out, int rounds)
{
    const int tid = threadIdx.x;
    uint2* p = buf + (size_t)blockIdx.x * 65536 + tid;
    long long total = 0; uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
        __syncthreads();
        const long long t0 = clock64();
        if (MODE == 0) { for (int i = 0; i < N; ++i) st_strong(p + i * blockDim.x, r, i); }
        if (MODE == 1) { for (int i = 0; i < N; ++i) st_weak(p + i * blockDim.x, r, i); }
        if (MODE == 2) { uint2 v[N]; for (int i = 0; i < N; ++i) v[i] = ld_strong(p + i * blockDim.x); for (int i = 0; i < N; ++i) acc += v[i].x + v[i].y; }
        if (MODE == 3) { uint2 v[N]; for (int i = 0; i < N; ++i) v[i] = ld_weak(p + i * blockDim.x); for (int i = 0; i < N; ++i) acc += v[i].x + v[i].y; }
        if (MODE == 4) { for (int i = 0; i < N; i += 2) st_strong4((uint4*)(p + i * blockDim.x) + tid, r, i); }
        if (MODE == 5) { for (int i = 0; i < N; ++i) st_strong(p + i * blockDim.x, r, i); __threadfence(); }
        const long long t1 = clock64();
        total += t1 - t0;
    }
    if (tid == 0) { out[blockIdx.x] = total / rounds; if (acc == 0x12345) out[blockIdx.x] = 0; }
}
int main()
{
    uint2* buf; long long* out; cudaMalloc(&buf, 148ull * 65536 * 32); cudaMemset(buf, 0, 148ull * 65536 * 32); cudaMallocManaged(&out, 148 * 8);
    const char* names[] = {"st strong v2", "st weak cg v2", "ld strong v2", "ld weak cg v2", "st strong v4 (half as many)", "st strong v2 + threadfence"};
#define RUN(M, N, T) k<M, N><<<148, T>>>(buf, out, 200); cudaDeviceSynchronize(); printf("%-28s N=%d threads=%4d: %lld cycles (CTA 0), %lld (CTA 100)  err=%d\n", names[M], N, T, out[0], out[100], (int)cudaGetLastError());
    for (int T : {32, 1024}) {
        RUN(0, 1, T) RUN(0, 4, T) RUN(0, 8, T) RUN(1, 4, T) RUN(1, 8, T) RUN(2, 1, T) RUN(2, 4, T) RUN(2, 8, T) RUN(3, 4, T) RUN(3, 8, T) RUN(4, 8, T) RUN(5, 4, T)
    }
    return 0;
}
