// kern_f_ep_tv.cu -- one instantiation of the fused scanline kernel (see scanline_kernels.cuh).
// R = float; <VHS, chroma delay, full output lowpass> = <true, 14, false>.
#include "scanline_kernels.cuh"
namespace cvs {
CVS_DEFINE_LAUNCH_FIELDS(float, true, 14, false)
}
