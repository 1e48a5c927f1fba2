// jb_single_light.cu -- single-plugin kernels of Infer, Punch, Saturator, Width, Cohere (fast math) and the dispatcher
#include "jb_single.cuh"

#include <cuda.h> // CUtensorMap and the cuTensorMapEncodeTiled prototype; the entry point itself is resolved at run time

extern "C" int jbk_single_exact(const ProcArgs* args, int grid, void* stream);
extern "C" int jbk_single_texture_a(const ProcArgs* args, int grid, void* stream);
extern "C" int jbk_single_texture_b(const ProcArgs* args, int grid, void* stream);
extern "C" int jbk_single_motion(const ProcArgs* args, int grid, void* stream);

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encodeTiled()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return (EncodeTiledFn) p;
    }();
    return fn;
}

// The launch's input rows as a 3-D tensor {sample, channel, clip} of fp32, box {16, 1, 32}, 64-byte swizzle (jb_lane.cuh).
bool fillTensorMap(ProcArgs& a)
{
    static_assert(sizeof(CUtensorMap) == sizeof(a.tmapIn), "ProcArgs::tmapIn holds one CUtensorMap");
    EncodeTiledFn enc = encodeTiled();
    if (enc == nullptr || a.nCh != 2 || a.clipMap != nullptr || (a.rowPitch & 3) != 0)
        return false;
    const cuuint64_t dims[3] = { (cuuint64_t) a.nSamples, 2, (cuuint64_t) a.nClips };
    const cuuint64_t strides[2] = { (cuuint64_t) a.rowPitch * 4, (cuuint64_t) a.rowPitch * 8 };
    const cuuint32_t box[3] = { TMA_S, 1, 32 }, elem[3] = { 1, 1, 1 };
    return enc(reinterpret_cast<CUtensorMap*>(a.tmapIn), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(a.in), dims, strides, box, elem,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace

// Launch the single-plugin kernel of args->slot[0] (args->chainLen == 1).  Returns a cudaError_t.
extern "C" int jbk_launch_single(const ProcArgs* argsIn, int grid, void* stream)
{
    cudaStream_t st = (cudaStream_t) stream;
    ProcArgs local;
    const ProcArgs* args = argsIn;
    if (argsIn->octets == 3) { // TMA tile streaming needs the tensor map of this launch's input; without one: cp.async rings
        local = *argsIn;
        if (!fillTensorMap(local))
            local.octets = 1;
        args = &local;
    }
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
