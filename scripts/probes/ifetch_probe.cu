// ifetch_probe.cu -- instruction supply on B200: issue rate of independent integer work as a function of the loop
// body size, with (a) every warp of a scheduler in the SAME loop and (b) the four warps of a scheduler in four
// DIFFERENT loops (what a role-per-warp kernel does when the roles are rotated over the schedulers).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ifetch_probe ifetch_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// N groups of 8 independent LOP3+IADD3 pairs = 16 N instructions, 256 N bytes
template <int N, uint32_t SALT>
__device__ __forceinline__ void body(uint32_t (&v)[8], uint32_t m) {
#pragma unroll
    for (int g = 0; g < N; g++) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (v[j] ^ m) + (SALT + (uint32_t)(g * 8 + j) * 2654435761u);
    }
}

template <int N, int MODE>   // MODE 0: one loop for all warps; 1: loop = warp % 4 (one per scheduler); 2: loop = (warp / 4) % 4 (four per scheduler)
__global__ void __launch_bounds__(512, 1) k(uint32_t *out, uint32_t m, int iters, long long *cyc) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = threadIdx.x * 40503u + j;
    const int warp = threadIdx.x >> 5;
    const int which = MODE == 0 ? 0 : (MODE == 1 ? (warp & 3) : ((warp >> 2) & 3));
    __syncthreads();
    const long long t0 = clock64();
    if (which == 0) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) body<N, 0x11111111u>(v, m);
    } else if (which == 1) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) body<N, 0x22222222u>(v, m);
    } else if (which == 2) {
#pragma unroll 1
        for (int i = 0; i < iters; i++) body<N, 0x33333333u>(v, m);
    } else {
#pragma unroll 1
        for (int i = 0; i < iters; i++) body<N, 0x44444444u>(v, m);
    }
    const long long t1 = clock64();
    uint32_t u = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) u ^= v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = u;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
}

template <int N, int MODE>
void run(int warps_per_smsp, uint32_t *out, long long *cyc) {
    const int nt = 128 * warps_per_smsp, ctas = 148;
    const int iters = 200000 / N > 50 ? 200000 / N : 50;
    for (int rep = 0; rep < 2; rep++) k<N, MODE><<<ctas, nt>>>(out, 0x5bd1e995u, iters, cyc);
    cudaDeviceSynchronize();
    const int nw = ctas * 16;
    long long *h = new long long[nw];
    cudaMemcpy(h, cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    int cnt = 0;
    for (int c = 0; c < ctas; c++)
        for (int w = 0; w < nt / 32; w++) { mean += (double)h[c * 16 + w]; cnt++; }
    mean /= cnt;
    delete[] h;
    const double instr = (double)iters * N * 16;
    printf("body %6.1f KB (%5d instr), %d warps/scheduler, %s: %.3f instr/cycle/scheduler\n", N * 256 / 1024.0, N * 16, warps_per_smsp,
           MODE == 0 ? "same loop      " : (MODE == 1 ? "loop per sched " : "4 loops / sched"), instr * warps_per_smsp / mean);
}

template <int N>
void sweep(uint32_t *out, long long *cyc) {
    run<N, 0>(2, out, cyc);
    run<N, 0>(4, out, cyc);
    run<N, 1>(4, out, cyc);
    run<N, 2>(4, out, cyc);
}

int main() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(uint32_t));
    cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
    sweep<8>(out, cyc);      // 2 KB
    sweep<16>(out, cyc);     // 4 KB
    sweep<24>(out, cyc);     // 6 KB
    sweep<32>(out, cyc);     // 8 KB
    sweep<48>(out, cyc);     // 12 KB
    sweep<64>(out, cyc);     // 16 KB
    sweep<96>(out, cyc);     // 24 KB
    sweep<128>(out, cyc);    // 32 KB
    sweep<192>(out, cyc);    // 48 KB
    sweep<256>(out, cyc);    // 64 KB
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
