// Micro-benchmarks of instruction latencies that matter for the persistent CG kernel's sync phases (diagnostics, not product code).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void k(double* out, unsigned* gcounter, double* gvals, int n) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 0.5;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    long long c0, c1;
    // (a) dependent DADD chain
    double v = threadIdx.x;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { v += 1.25; v += 2.5; v += 0.75; v += 1.5; }
    c1 = clock64();
    if (threadIdx.x == 0) out[0] = double(c1 - c0) / (4.0 * n);
    // (b) dependent FADD chain
    float f = threadIdx.x;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { f += 1.25f; f += 2.5f; f += 0.75f; f += 1.5f; }
    c1 = clock64();
    if (threadIdx.x == 0) out[1] = double(c1 - c0) / (4.0 * n);
    // (c) independent DADDs (throughput, one warp)
    double a0 = v, a1 = v + 1, a2 = v + 2, a3 = v + 3, a4 = v + 4, a5 = v + 5, a6 = v + 6, a7 = v + 7;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { a0 += 1.25; a1 += 1.25; a2 += 1.25; a3 += 1.25; a4 += 1.25; a5 += 1.25; a6 += 1.25; a7 += 1.25; }
    c1 = clock64();
    if (threadIdx.x == 0) out[2] = double(c1 - c0) / (8.0 * n);
    v = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    // (d) globaltimer read, back to back
    unsigned long long t = 0;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) t += gt();
    c1 = clock64();
    if (threadIdx.x == 0) out[3] = double(c1 - c0) / 64.0;
    // (e) LDS.64 dependent chain
    int idx = threadIdx.x;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) { double x = sm[idx & 1023]; idx = int(x) + i; }
    c1 = clock64();
    if (threadIdx.x == 0) out[4] = double(c1 - c0) / n;
    // (f) double shuffle + DADD tree of 4 values (what acc_warp_sum does)
    double q0 = v, q1 = v * 2, q2 = v * 3, q3 = v * 4;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { q0 += __shfl_down_sync(0xffffffffu, q0, o); q1 += __shfl_down_sync(0xffffffffu, q1, o); q2 += __shfl_down_sync(0xffffffffu, q2, o); q3 += __shfl_down_sync(0xffffffffu, q3, o); }
    }
    c1 = clock64();
    if (threadIdx.x == 0) out[5] = double(c1 - c0) / 64.0;
    // (g) acquire poll on a counter that is already at its target
    c0 = clock64();
    unsigned s = 0;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) { unsigned x; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(x) : "l"(gcounter) : "memory"); s += x; }
    c1 = clock64();
    if (threadIdx.x == 0) out[6] = double(c1 - c0) / 64.0;
    // (h) release RMW (membar + red) with nothing outstanding
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(gcounter + 32), "r"(1u) : "memory"); }
    c1 = clock64();
    if (threadIdx.x == 0) out[7] = double(c1 - c0) / 64.0;
    // (i) L2 load round trip (ld.cg of a line another SM never touched), dependent chain
    c0 = clock64();
    double w = 0; int j = threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 64; ++i) { w += __ldcg(gvals + j); j = (j + 1024 + int(w)) & 0xFFFFF; }
    c1 = clock64();
    if (threadIdx.x == 0) out[8] = double(c1 - c0) / 64.0;
    if (threadIdx.x == 0) { out[9] = v + f + double(t) + idx + q0 + q1 + q2 + q3 + s + w; }
}
int main() {
    double* out; unsigned* cnt; double* vals;
    cudaMalloc(&out, 16 * sizeof(double)); cudaMalloc(&cnt, 1024); cudaMemset(cnt, 0, 1024);
    cudaMalloc(&vals, (1 << 20) * sizeof(double)); cudaMemset(vals, 0, (1 << 20) * sizeof(double));
    for (int rep = 0; rep < 2; ++rep) k<<<1, 128>>>(out, cnt, vals, 256);
    double h[16];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[] = {"DADD dependent latency", "FADD dependent latency", "DADD issue interval (8 independent, 1 warp)", "globaltimer read", "LDS.64 dependent", "4x double warp-shuffle tree",
                           "ld.acquire.gpu poll trip", "red.release.gpu trip", "ld.cg L2 round trip"};
    for (int i = 0; i < 9; ++i) printf("%-48s %8.1f cycles\n", names[i], h[i]);
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
