// kern_f_comp_full.cu -- one instantiation of the fused scanline kernel (see scanline_kernels.cuh).
// R = float; <VHS, chroma delay, full output lowpass> = <false, 9, true>.
#include "scanline_kernels.cuh"
namespace cvs {
CVS_DEFINE_LAUNCH_FIELDS(float, false, 9, true)
}
