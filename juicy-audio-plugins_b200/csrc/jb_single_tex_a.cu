// jb_single_tex_a.cu -- Texture gel, metal, flesh
#include "jb_single.cuh"

extern "C" int jbk_single_texture_a(const ProcArgs* args, int grid, void* stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    switch (args->slot[0].c.tex.material) {
        case 0: return (int) launch_single<MainTexture<0>, PreAna>(*args, grid, st);
        case 1: return (int) launch_single<MainTexture<1>, PreAna>(*args, grid, st);
        default: return (int) launch_single<MainTexture<4>, PreAna>(*args, grid, st);
    }
}
