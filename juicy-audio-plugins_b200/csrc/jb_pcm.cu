// jb_pcm.cu -- 16-bit PCM <-> fp32 on the device, for jb_process_host_pcm16 (the step either side of the
// render is host<->device streaming over PCIe, an order of magnitude slower than HBM; 16-bit sources need not cross it as
// 32-bit floats).  The conversion rule is jb_wav.cpp's, i.e. what a DAW's file reader / writer does around the plugin:
// reading v = s / 32768, writing s = round-half-even(v * 32768) limited to +-32767.  Both directions are exact restatements:
// the scaling is by a power of two, so float and double arithmetic round alike.
#include "jb_kernels.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// rows x n samples; srcPitch / dstPitch in elements.  One thread converts eight samples (16 B in, 32 B out) when aligned.
__global__ void jb_pcm16_to_float_kernel(const int16_t* __restrict__ src, float* __restrict__ dst, long long rows, int n,
                                         long long srcPitch, long long dstPitch)
{
    const int n8 = n >> 3;
    const long long total = rows * (long long) (n8 + 1);
    const long long stride = (long long) gridDim.x * blockDim.x;
    const bool vec = ((srcPitch | dstPitch) & 7) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15u) | (reinterpret_cast<uintptr_t>(dst) & 31u)) == 0;
    for (long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const long long row = idx / (n8 + 1);
        const int o = (int) (idx - row * (n8 + 1));
        const int16_t* s = src + row * srcPitch;
        float* d = dst + row * dstPitch;
        if (o < n8) {
            if (vec) {
                const int4 v = *reinterpret_cast<const int4*>(s + 8 * o);
                const int w[4] = { v.x, v.y, v.z, v.w };
                float f[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    f[2 * k] = (float) (int16_t) (w[k] & 0xffff) * (1.0f / 32768.0f);
                    f[2 * k + 1] = (float) (int16_t) (w[k] >> 16) * (1.0f / 32768.0f);
                }
                *reinterpret_cast<float4*>(d + 8 * o) = make_float4(f[0], f[1], f[2], f[3]);
                *reinterpret_cast<float4*>(d + 8 * o + 4) = make_float4(f[4], f[5], f[6], f[7]);
            } else {
                for (int k = 0; k < 8; ++k)
                    d[8 * o + k] = (float) s[8 * o + k] * (1.0f / 32768.0f);
            }
        } else {
            for (int i = 8 * n8; i < n; ++i)
                d[i] = (float) s[i] * (1.0f / 32768.0f);
        }
    }
}

__device__ __forceinline__ int quantize16(float v)
{
    const int q = __float2int_rn(v * 32768.0f); // nearbyint in the default rounding mode: half to even; saturates for huge |v|
    return max(-32767, min(32767, q));
}

__global__ void jb_float_to_pcm16_kernel(const float* __restrict__ src, int16_t* __restrict__ dst, long long rows, int n,
                                         long long srcPitch, long long dstPitch)
{
    const int n8 = n >> 3;
    const long long total = rows * (long long) (n8 + 1);
    const long long stride = (long long) gridDim.x * blockDim.x;
    const bool vec = ((srcPitch | dstPitch) & 7) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15u) | (reinterpret_cast<uintptr_t>(src) & 31u)) == 0;
    for (long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const long long row = idx / (n8 + 1);
        const int o = (int) (idx - row * (n8 + 1));
        const float* s = src + row * srcPitch;
        int16_t* d = dst + row * dstPitch;
        if (o < n8) {
            if (vec) {
                const float4 a = *reinterpret_cast<const float4*>(s + 8 * o), b = *reinterpret_cast<const float4*>(s + 8 * o + 4);
                const float f[8] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w };
                int w[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    w[k] = (quantize16(f[2 * k]) & 0xffff) | (quantize16(f[2 * k + 1]) << 16);
                *reinterpret_cast<int4*>(d + 8 * o) = make_int4(w[0], w[1], w[2], w[3]);
            } else {
                for (int k = 0; k < 8; ++k)
                    d[8 * o + k] = (int16_t) quantize16(s[8 * o + k]);
            }
        } else {
            for (int i = 8 * n8; i < n; ++i)
                d[i] = (int16_t) quantize16(s[i]);
        }
    }
}

int gridFor(long long rows, int n)
{
    const long long work = rows * (long long) ((n >> 3) + 1);
    long long blocks = (work + 255) / 256;
    if (blocks > 148 * 16)
        blocks = 148 * 16;
    return (int) (blocks < 1 ? 1 : blocks);
}

} // namespace

extern "C" int jbk_launch_pcm16_to_float(const int16_t* src, float* dst, long long rows, int n, long long srcPitch, long long dstPitch,
                                         void* stream)
{
    if (rows <= 0 || n <= 0)
        return 0;
    jb_pcm16_to_float_kernel<<<gridFor(rows, n), 256, 0, (cudaStream_t) stream>>>(src, dst, rows, n, srcPitch, dstPitch);
    jbk_note_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int jbk_launch_float_to_pcm16(const float* src, int16_t* dst, long long rows, int n, long long srcPitch, long long dstPitch,
                                         void* stream)
{
    if (rows <= 0 || n <= 0)
        return 0;
    jb_float_to_pcm16_kernel<<<gridFor(rows, n), 256, 0, (cudaStream_t) stream>>>(src, dst, rows, n, srcPitch, dstPitch);
    jbk_note_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
