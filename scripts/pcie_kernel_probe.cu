// Probe: GPU-driven gather/scatter of the field rows over PCIe (mapped pinned host memory) vs copy engines.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); return 1;}}while(0)
// rows: n_pics * nl rows of `wbytes`; src row r of picture k at src + k*spic + (y0+2r)*sstride
__global__ void k_rows(uint4* __restrict__ dst, size_t dpic, size_t dstride, const uint4* __restrict__ src, size_t spic, size_t sstride,
                       int nl, int y0, int w16, int npics) {
    const int total_rows = nl * npics;
    for (int row = blockIdx.x; row < total_rows; row += gridDim.x) {
        const int k = row / nl, r = row - k * nl;
        const uint4* s = (const uint4*)((const char*)src + k * spic + (size_t)(y0 + 2 * r) * sstride);
        uint4* d = (uint4*)((char*)dst + k * dpic + (size_t)(y0 + 2 * r) * dstride);
        for (int i = threadIdx.x; i < w16; i += blockDim.x) d[i] = s[i];
    }
}
int main() {
    const int W = 1920, H = 1080, N = 128, nl = 540; const size_t stride = W * 4, pic = stride * H;
    char *h_in, *h_out, *d_in, *d_out;
    CK(cudaHostAlloc(&h_in, pic * N, cudaHostAllocDefault)); CK(cudaHostAlloc(&h_out, pic * N, cudaHostAllocDefault));
    CK(cudaMalloc(&d_in, pic * N)); CK(cudaMalloc(&d_out, pic * N));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double gb = (double)N * nl * stride / 1e9;
    for (int ctas : {16, 32, 64, 128}) for (int mode = 0; mode < 3; mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, 0));
            if (mode == 0 || mode == 2) k_rows<<<ctas, 256, 0, s1>>>((uint4*)d_in, pic, stride, (const uint4*)h_in, pic, stride, nl, 1, W / 4, N);
            if (mode == 1 || mode == 2) k_rows<<<ctas, 256, 0, s2>>>((uint4*)h_out, pic, stride, (const uint4*)d_out, pic, stride, nl, 1, W / 4, N);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e1, 0)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("kernel ctas=%3d %s: %.1f GB/s%s\n", ctas, mode == 0 ? "H2D " : mode == 1 ? "D2H " : "both", gb / (best / 1e3), mode == 2 ? " each way" : "");
    }
    // copy engines, 2D strided, per picture (what run_host does today)
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0, 0));
            for (int k = 0; k < N; k++) {
                if (mode != 1) CK(cudaMemcpy2DAsync(d_in + k * pic + stride, 2 * stride, h_in + k * pic + stride, 2 * stride, stride, nl, cudaMemcpyHostToDevice, s1));
                if (mode != 0) CK(cudaMemcpy2DAsync(h_out + k * pic + stride, 2 * stride, d_out + k * pic + stride, 2 * stride, stride, nl, cudaMemcpyDeviceToHost, s2));
            }
            CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e1, 0)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("memcpy2D %s: %.1f GB/s%s\n", mode == 0 ? "H2D " : mode == 1 ? "D2H " : "both", gb / (best / 1e3), mode == 2 ? " each way" : "");
    }
    return 0;
}
