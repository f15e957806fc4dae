// fp64_mix_probe.cu -- does an FP64 instruction keep a B200 scheduler's dispatch port busy for its two pipe cycles?
// Per iteration every warp issues NF independent DADDs and NI independent integer LOP3/IADD3 instructions.
// If dispatch is blocked: cycles per iteration per scheduler ~ W * (2 NF + NI); if not: W * max(2 NF, NF + NI).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_mix_probe fp64_mix_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NF, int NI, int KIND>
__global__ void k(double *out, uint32_t *iout, double a, uint32_t m, int iters, long long *cyc) {
    double x[NF > 0 ? NF : 1];
    uint32_t v[NI > 0 ? NI : 1];
#pragma unroll
    for (int j = 0; j < NF; j++) x[j] = 1.0 + threadIdx.x * 1e-3 + j;
#pragma unroll
    for (int j = 0; j < NI; j++) v[j] = threadIdx.x * 2654435761u + j;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int j = 0; j < (NF > NI ? NF : NI); j++) {
                if (j < NF) x[j] = __dadd_rn(x[j], a);
                if (j < NI) {
                    if (KIND == 0) v[j] = (v[j] ^ m) + 0x9E3779B9u;            // LOP3 + IADD3 (ALU)  -> counts as 2
                    else if (KIND == 1) v[j] = v[j] * m + 12345u;               // IMAD (FMA pipe)
                    else v[j] = __float_as_uint(__fmaf_rn(__uint_as_float(v[j]), 1.0001f, 0.5f));   // FFMA
                }
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    uint32_t u = 0;
#pragma unroll
    for (int j = 0; j < NF; j++) s += x[j];
#pragma unroll
    for (int j = 0; j < NI; j++) u ^= v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = u;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int NF, int NI, int KIND>
void run(int warps_per_smsp, double *out, uint32_t *iout, long long *cyc) {
    const int iters = 2000, nt = 128 * warps_per_smsp, ctas = 148;
    for (int rep = 0; rep < 2; rep++) k<NF, NI, KIND><<<ctas, nt>>>(out, iout, 0.37, 0x5bd1e995u, iters, cyc);
    cudaDeviceSynchronize();
    const int nw = ctas * nt / 32;
    long long *h = new long long[nw];
    cudaMemcpy(h, cyc, nw * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < nw; i++) mean += (double)h[i];
    mean /= nw;
    delete[] h;
    const int ni_instr = NI * (KIND == 0 ? 2 : 1);
    const double per_iter = mean / (iters * 4.0);
    printf("NF %d DADD + %2d %s, %d warps/scheduler: %6.2f cycles per group per scheduler-warp, %6.2f per scheduler;  blocked model %d, free model %d\n",
           NF, ni_instr, KIND == 0 ? "ALU ops " : (KIND == 1 ? "IMAD    " : "FFMA    "), warps_per_smsp, per_iter, per_iter / 1.0,
           warps_per_smsp * (2 * NF + ni_instr), warps_per_smsp * ((2 * NF > NF + ni_instr) ? 2 * NF : NF + ni_instr));
}

int main() {
    double *out; uint32_t *iout; long long *cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(double));
    cudaMalloc(&iout, 148 * 512 * sizeof(uint32_t));
    cudaMalloc(&cyc, 148 * 16 * sizeof(long long));
    for (int w = 2; w <= 4; w += 2) {
        run<8, 0, 0>(w, out, iout, cyc);
        run<0, 8, 0>(w, out, iout, cyc);
        run<8, 4, 0>(w, out, iout, cyc);
        run<8, 8, 0>(w, out, iout, cyc);
        run<4, 8, 0>(w, out, iout, cyc);
        run<8, 8, 1>(w, out, iout, cyc);
        run<8, 16, 1>(w, out, iout, cyc);
        run<8, 8, 2>(w, out, iout, cyc);
        run<8, 16, 2>(w, out, iout, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
