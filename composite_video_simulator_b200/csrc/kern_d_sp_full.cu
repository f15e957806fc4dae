// kern_d_sp_full.cu -- one instantiation of the fused scanline kernel (see scanline_kernels.cuh).
// R = double; <VHS, chroma delay, full output lowpass> = <true, 9, true>.
#include "scanline_kernels.cuh"
namespace cvs {
CVS_DEFINE_LAUNCH_FIELDS(double, true, 9, true)
}
