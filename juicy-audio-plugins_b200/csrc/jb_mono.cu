// jb_mono.cu -- the generic lane kernel instantiated for one-channel buses (see `MONO` in jb_lane.cuh), fast math.  Its own
// translation unit so that it compiles beside the stereo one; the exact-math instantiation lives in jb_mono_exact.cu.
#include "jb_lane.cuh"

extern "C" int jbk_launch_mono_exact(const ProcArgs* args, int grid, void* stream);

extern "C" int jbk_launch_mono(const ProcArgs* args, int grid, void* stream)
{
    // exact math (JB_MATH_EXACT, or JB_MATH_AUTO with a shaper in front of another plugin): the instantiation whose
    // Saturator / Punch call the C library's own tanh / pow, so a mono chain honours the math mode like a stereo one
    if (args->exactMath)
        return jbk_launch_mono_exact(args, grid, stream);
    jb_process_kernel<true, false><<<grid, JB_CTA_THREADS, lane_smem_bytes(0), (cudaStream_t) stream>>>(*args);
    return (int) cudaGetLastError();
}
