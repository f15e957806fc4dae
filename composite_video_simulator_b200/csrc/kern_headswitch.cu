// kern_headswitch.cu -- launchers of the head-switch pre-pass (see scanline_kernels.cuh).
#include "scanline_kernels.cuh"
namespace cvs {
template <typename R>
cudaError_t launch_headswitch(const LaunchArgs<R> &a, const HsItem *items, int nitems, cudaStream_t st) {
    const int ctas = (nitems + kHsNT - 1) / kHsNT;
    k_headswitch<R><<<ctas, kHsNT, 0, st>>>(a, items, nitems);
    return cudaGetLastError();
}
template cudaError_t launch_headswitch<float>(const LaunchArgs<float> &, const HsItem *, int, cudaStream_t);
template cudaError_t launch_headswitch<double>(const LaunchArgs<double> &, const HsItem *, int, cudaStream_t);
}
