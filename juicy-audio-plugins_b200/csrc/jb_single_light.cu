// jb_single_light.cu -- single-plugin kernels of Infer, Punch, Saturator, Width, Cohere (fast math) and the dispatcher
#include "jb_single.cuh"

extern "C" int jbk_single_exact(const ProcArgs* args, int grid, void* stream);
extern "C" int jbk_single_texture_a(const ProcArgs* args, int grid, void* stream);
extern "C" int jbk_single_texture_b(const ProcArgs* args, int grid, void* stream);
extern "C" int jbk_single_motion(const ProcArgs* args, int grid, void* stream);

// Launch the single-plugin kernel of args->slot[0] (args->chainLen == 1).  Returns a cudaError_t.
extern "C" int jbk_launch_single(const ProcArgs* args, int grid, void* stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    const SlotDesc& d = args->slot[0];
    switch (d.kind) {
        case K_INFER: return (int) launch_single<MainInfer, PreAna>(*args, grid, st);
        case K_PUNCH: return args->exactMath ? jbk_single_exact(args, grid, stream) : (int) launch_single<MainPunch<false>, PreAna>(*args, grid, st);
        case K_SAT: return args->exactMath ? jbk_single_exact(args, grid, stream) : (int) launch_single<MainSat<false>, PreAna>(*args, grid, st);
        case K_WIDTH: return (int) launch_single<MainWidth, PreAna>(*args, grid, st);
        case K_COHERE: return (int) launch_single<MainCohere, PreCohere>(*args, grid, st);
        case K_MOTION: return jbk_single_motion(args, grid, stream);
        case K_TEXTURE: {
            const int m = d.c.tex.material;
            return (m == 2 || m == 3) ? jbk_single_texture_b(args, grid, stream) : jbk_single_texture_a(args, grid, stream);
        }
        default: return (int) cudaErrorInvalidValue;
    }
}
