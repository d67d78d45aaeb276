// Does warp divergence persist across BAR.SYNC / loops, and what does it cost a shuffle tree? (diagnostics)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __noinline__ bool poll(const unsigned* c, unsigned target) {
    const long long t0 = clock64();
    while (ld_acq(c) < target) { if (clock64() - t0 > 100000) return false; }
    return true;
}
__device__ __forceinline__ long long tree(double& q0, double& q1, double& q2, double& q3) {
    const long long c0 = clock64();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { q0 += __shfl_down_sync(0xffffffffu, q0, o); q1 += __shfl_down_sync(0xffffffffu, q1, o); q2 += __shfl_down_sync(0xffffffffu, q2, o); q3 += __shfl_down_sync(0xffffffffu, q3, o); }
    return clock64() - c0;
}
__global__ void k(double* out, unsigned* cnt, int variant, int n) {
    __shared__ double sm[256];
    __shared__ int flag;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = i;
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    double q0 = threadIdx.x, q1 = 2, q2 = 3, q3 = 4;
    long long cyc = 0;
    if (variant == 1) {          // lane 0 polls (loop with two exits), then a CTA barrier
        if (threadIdx.x == 0 && !poll(cnt, 0u)) flag = 1;
        __syncthreads();
    } else if (variant == 2) {   // same + __syncwarp
        if (threadIdx.x == 0 && !poll(cnt, 0u)) flag = 1;
        __syncthreads();
        __syncwarp();
    } else if (variant == 3) {   // lane-dependent trip counts
        for (int u = threadIdx.x; u < n; u += 32) q0 += sm[u];
    } else if (variant == 4) {   // lane-dependent trip counts + syncwarp
        for (int u = threadIdx.x; u < n; u += 32) q0 += sm[u];
        __syncwarp();
    } else if (variant == 5) {   // named barrier after a one-lane section
        if (threadIdx.x == 0 && !poll(cnt, 0u)) flag = 1;
        asm volatile("bar.sync 1, 128;" ::: "memory");
    } else if (variant == 6) {   // polling lane is NOT made to leave early (no second exit)
        if (threadIdx.x == 0) { while (ld_acq(cnt) < 0u) { } }
        __syncthreads();
    }
    if (threadIdx.x < 32) {
        cyc = tree(q0, q1, q2, q3);
        if (threadIdx.x == 0) { out[variant] = double(cyc); out[16 + variant] = q0 + q1 + q2 + q3 + flag; }
    }
}
int main() {
    double* out; unsigned* cnt;
    cudaMalloc(&out, 64 * sizeof(double)); cudaMalloc(&cnt, 1024); cudaMemset(cnt, 0, 1024);
    for (int v = 0; v <= 6; ++v) { k<<<1, 128>>>(out, cnt, v, 13); k<<<1, 128>>>(out, cnt, v, 13); }
    double h[64];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"converged", "lane-0 poll (2 exits) + __syncthreads", "... + __syncwarp", "lane-dependent loop", "... + __syncwarp", "lane-0 poll + named barrier", "lane-0 poll (1 exit) + __syncthreads"};
    for (int i = 0; i <= 6; ++i) printf("%-48s %8.0f cycles\n", names[i], h[i]);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
