// yuv422_kernels.cu -- instantiation and launchers of the 4:2:2 kernels (yuv422_kernels.cuh).
#define CVS_YUV422_DEFINE_KERNELS
#include "yuv422_kernels.cuh"

#include <cstdlib>

namespace cvs422 {

// CVS422_GENERAL=1 forces the general kernel (read at every launch, so a test can switch it inside one process:
// both kernels must give the same pictures)
static bool getenv_general() {
    const char *e = std::getenv("CVS422_GENERAL");
    return e && e[0] == '1';
}

cudaError_t launch_yuv422(const Launch422 &a, const HsItem422 *d_items, int nitems, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_yuv422, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem422::total_max);
    if (e != cudaSuccess) return e;
    if (a.packed ? a.total_warps > 1 : a.warps_per_field > 1) {
        k_yuv422_halo<<<a.packed ? a.total_warps : a.nfields * a.warps_per_field, 128, 0, st>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (nitems > 0) {
        k_yuv422_headswitch<<<(nitems + kHsNT - 1) / kHsNT, kHsNT, 0, st>>>(a, d_items, nitems);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    // the fast kernel where it applies (aligned planes, whole blocks, common switches, no pre-pass rows)
    if (fast_row_ok(a.K) && a.vec && nitems == 0 && !getenv_general()) k_yuv422_fast<<<a.total_warps, kNT, Smem422::total(a.K), st>>>(a);
    else k_yuv422<<<a.total_warps, kNT, Smem422::total(a.K), st>>>(a);
    return cudaGetLastError();
}

cudaError_t occupancy_yuv422(const K422 &K, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(k_yuv422, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem422::total_max);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_yuv422, kNT, Smem422::total(K));
}

cudaError_t launch_render_field(const RenderArgs &a, cudaStream_t st) {
    int maxb = a.row_bytes[0];
    for (int p = 1; p < 3; p++) maxb = a.row_bytes[p] > maxb ? a.row_bytes[p] : maxb;
    const int rows = (a.dst_h > a.field) ? (a.dst_h - a.field + 1) / 2 : 0;
    if (rows <= 0 || maxb <= 0) return cudaSuccess;
    dim3 grid((unsigned)((maxb + 255) / 256), (unsigned)rows, 3);
    k_render_field<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace cvs422
