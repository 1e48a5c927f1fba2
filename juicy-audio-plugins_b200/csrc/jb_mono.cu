// jb_mono.cu -- the generic lane kernel instantiated for one-channel buses (see `MONO` in jb_lane.cuh).  Its own
// translation unit so that it compiles beside the stereo one.
#include "jb_lane.cuh"

extern "C" int jbk_launch_mono(const ProcArgs* args, int grid, void* stream)
{
    jb_process_kernel<true><<<grid, JB_CTA_THREADS, lane_smem_bytes(0), (cudaStream_t) stream>>>(*args);
    return (int) cudaGetLastError();
}
