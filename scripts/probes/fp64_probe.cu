// fp64_probe.cu -- B200 micro-measurements behind the 4:2:2 kernel's design (profiles/fp64_probe_r2.txt):
// latency of dependent FP64 operations and how many independent chains x warps per scheduler saturate the FP64 pipe;
// the same for the 64-bit conversions (I2F.F64 / F2I.F64) and the saturating pack.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { OP_DADD, OP_DMUL, OP_DFMA, OP_POLE, OP_CONV, OP_PACK, OP_MIX };

template <int OP>
__device__ __forceinline__ double op(double x, double a, double b) {
    if (OP == OP_DADD) return __dadd_rn(x, a);
    if (OP == OP_DMUL) return __dmul_rn(x, b);
    if (OP == OP_DFMA) return __fma_rn(x, b, a);
    if (OP == OP_POLE) {                       // p = s*a + (p - p*a): DMUL -> DADD -> DADD on the carried value
        const double t = __dmul_rn(a, b);
        const double u = __dsub_rn(x, __dmul_rn(x, b));
        return __dadd_rn(t, u);
    }
    if (OP == OP_CONV) return __dadd_rn(__uint2double_rn((uint32_t)__double2int_rz(x) & 0xFFu), a);   // F2I.F64 -> LOP -> I2F.F64 -> DADD
    if (OP == OP_PACK) {
        uint32_t d, v = (uint32_t)__double_as_longlong(x);
        asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v), "r"(v), "r"(0));
        return __longlong_as_double((long long)d | 0x4000000000000000ll);
    }
    return x;
}

template <int OP, int ILP>
__global__ void k(double *out, double a, double b, int iters, long long *cyc) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = 1.0 + threadIdx.x * 1e-3 + j;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int j = 0; j < ILP; j++) x[j] = op<OP>(x[j], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int OP, int ILP>
void run(const char *name, int warps_per_smsp, double *out, long long *cyc) {
    const int iters = 2000, nt = 128 * warps_per_smsp, ctas = 148;
    k<OP, ILP><<<ctas, nt>>>(out, 0.37, 0.999, iters, cyc);
    k<OP, ILP><<<ctas, nt>>>(out, 0.37, 0.999, iters, cyc);
    cudaDeviceSynchronize();
    const int nw = ctas * nt / 32;
    long long *h = new long long[nw];
    cudaMemcpy(h, cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < nw; i++) mean += (double)h[i];
    mean /= nw;
    delete[] h;
    const double ops = (double)iters * 8 * ILP;           // calls of op() per warp
    printf("%-5s ILP %d warps/SMSP %d: %7.2f cycles per dependent step, %6.3f op() per cycle per SMSP\n", name, ILP, warps_per_smsp,
           mean / (iters * 8.0), ops * warps_per_smsp / mean);
}

template <int OP>
void sweep(const char *name, double *out, long long *cyc) {
    for (int w = 1; w <= 4; w++) {
        run<OP, 1>(name, w, out, cyc);
        run<OP, 2>(name, w, out, cyc);
        run<OP, 4>(name, w, out, cyc);
        run<OP, 8>(name, w, out, cyc);
    }
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(double));
    cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
    sweep<OP_DADD>("DADD", out, cyc);
    sweep<OP_DMUL>("DMUL", out, cyc);
    sweep<OP_DFMA>("DFMA", out, cyc);
    sweep<OP_POLE>("POLE", out, cyc);
    sweep<OP_CONV>("CONV", out, cyc);
    sweep<OP_PACK>("PACK", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
