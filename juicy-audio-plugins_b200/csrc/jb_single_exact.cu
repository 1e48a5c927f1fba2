// jb_single_exact.cu -- Saturator / Punch with the C library's own tanh / pow (jb_libm.h)
#include "jb_single.cuh"

extern "C" int jbk_single_exact(const ProcArgs* args, int grid, void* stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    if (args->slot[0].kind == K_PUNCH)
        return (int) launch_single<MainPunch<true>, PreAna>(*args, grid, st);
    return (int) launch_single<MainSat<true>, PreAna>(*args, grid, st);
}
