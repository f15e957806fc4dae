// kern_n_sp_full.cu -- one instantiation of the fused scanline kernel (see scanline_kernels.cuh), fast noise mode.
// R = float; <VHS, chroma delay, full output lowpass, fast noise> = <true, 9, true, true>.
#include "scanline_kernels.cuh"
namespace cvs {
CVS_DEFINE_LAUNCH_FIELDS_NF(float, true, 9, true, true)
}
