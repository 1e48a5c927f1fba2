// jb_mono_exact.cu -- one-channel buses with the C library's own tanh / pow in Saturator / Punch (see jb_mono.cu)
#include "jb_lane.cuh"

extern "C" int jbk_launch_mono_exact(const ProcArgs* args, int grid, void* stream)
{
    jb_process_kernel<true, true><<<grid, JB_CTA_THREADS, lane_smem_bytes(0), (cudaStream_t) stream>>>(*args);
    return (int) cudaGetLastError();
}
