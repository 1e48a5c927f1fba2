// jb_single_tex_b.cu -- Texture wood, plastic (waveguide materials) and Motion
#include "jb_single.cuh"

extern "C" int jbk_single_texture_b(const ProcArgs* args, int grid, void* stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    if (args->slot[0].c.tex.material == 2)
        return (int) launch_single<MainTexture<2>, PreAna>(*args, grid, st);
    return (int) launch_single<MainTexture<3>, PreAna>(*args, grid, st);
}

extern "C" int jbk_single_motion(const ProcArgs* args, int grid, void* stream)
{
    return (int) launch_single<MainMotion, PreMotion>(*args, grid, (cudaStream_t) stream);
}
