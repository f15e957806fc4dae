// kern_d_comp_tv.cu -- one instantiation of the fused scanline kernel (see scanline_kernels.cuh).
// R = double; <VHS, chroma delay, full output lowpass> = <false, 9, false>.
#include "scanline_kernels.cuh"
namespace cvs {
CVS_DEFINE_LAUNCH_FIELDS(double, false, 9, false)
}
