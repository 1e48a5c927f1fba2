// Micro-benchmark: cycles per sample of the analyzer's envelope/onset recurrence for ONE warp alone on an SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -ftz=true -o env_chain env_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
struct C { float aS, rS, aL, rL, omaS, omrS, omaL, omrL; int len; };
__device__ __forceinline__ float4 ldq(const float4* p) { return *p; }

template <int VARIANT>
__global__ void k(const float* x, int n, C c, float* out, long long* cyc)
{
    // 256 samples per lane staged in shared memory (row pitch 260 floats: conflict-free for 16-byte reads), walked n/256 times
    __shared__ float4 tile[32][65];
    for (int q = 0; q < 64; ++q)
        tile[threadIdx.x][q] = reinterpret_cast<const float4*>(x + (size_t) threadIdx.x * n)[q];
    __syncthreads();
    const float4* p = &tile[threadIdx.x][0];
    float s = 0.f, l = 0.f, tracc = 0.f; int cool = 0, onsets = 0;
    long long t0 = clock64();
    for (int q = 0; q < n / 4; ++q) {
        float4 v = ldq(p + (q & 63));
        float m[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = fabsf(m[j]);
            if (VARIANT == 0 || VARIANT == 2) { // select coefficients
                bool us = a > s; s = (us ? c.omaS : c.omrS) * a + (us ? c.aS : c.rS) * s;
                bool ul = a > l; l = (ul ? c.omaL : c.omrL) * a + (ul ? c.aL : c.rL) * l;
            } else { // both candidates
                float su = c.omaS * a + c.aS * s, sd = c.omrS * a + c.rS * s; s = a > s ? su : sd;
                float lu = c.omaL * a + c.aL * l, ld = c.omrL * a + c.rL * l; l = a > l ? lu : ld;
            }
            float tr = fmaxf(0.f, s - l);
            tracc += tr;
            if (VARIANT != 2) {
                cool = max(cool - 1, 0);
                bool on = (tr > 0.045f) & (cool <= 0);
                onsets += on; cool = on ? c.len : cool;
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = s + l + tracc + cool + onsets;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
    const int n = 65536;
    float* x; float* out; long long* cyc;
    cudaMalloc(&x, sizeof(float) * 32 * n); cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    float* h = new float[32 * n];
    unsigned r = 1; for (int i = 0; i < 32 * n; ++i) { r = r * 1664525u + 1013904223u; h[i] = ((r >> 8) & 0xffff) / 65536.f - 0.5f; }
    cudaMemcpy(x, h, sizeof(float) * 32 * n, cudaMemcpyHostToDevice);
    C c { 0.993f, 0.9993f, 0.9996f, 0.99993f, 0.007f, 0.0007f, 0.0004f, 0.00007f, 1680 };
    long long hc;
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 32>>>(x, n, c, out, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); printf("select-coef + onset : %.1f cycles/sample\n", (double) hc / n);
        k<1><<<1, 32>>>(x, n, c, out, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); printf("both-candidates + onset: %.1f cycles/sample\n", (double) hc / n);
        k<2><<<1, 32>>>(x, n, c, out, cyc); cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost); printf("select-coef, no onset : %.1f cycles/sample\n", (double) hc / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
