// Micro-benchmark 2: cycles per sample of the analyzer's two attack/release envelopes (no onset machine;
// transient sum + group max kept) for W warps on one SM, several formulations of the select.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -ftz=true -o env_chain2 env_chain2.cu
#include <cstdio>
#include <cuda_runtime.h>
struct C { float aS, rS, aL, rL, omaS, omrS, omaL, omrL; };

__device__ __forceinline__ unsigned gt_mask(float a, float e)
{
    unsigned m;
    asm("set.gt.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(a), "f"(e));
    return m;
}
__device__ __forceinline__ float bitsel(unsigned m, float x, float y) // m ? x : y, bitwise
{
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(m), "r"(__float_as_uint(x)), "r"(__float_as_uint(y)));
    return __uint_as_float(r);
}

// VARIANT 0: predicate select of coefficients (current kernel)   1: mask + lop3 select of coefficients
//         2: both candidates + fmaxf (not bit-exact near ties)    3: both candidates + mask/lop3 select of the result
//         4: one envelope per lane (lanes 0-15 short, 16-31 long), predicate select, tr via shuffle
template <int VARIANT>
__global__ void k(const float* x, int n, C c, float* out, long long* cyc)
{
    extern __shared__ float4 tile[]; // [warps*32][65]
    const int t = threadIdx.x;
    for (int q = 0; q < 64; ++q)
        tile[t * 65 + q] = reinterpret_cast<const float4*>(x + (size_t) (t & 31) * n)[q];
    __syncthreads();
    const float4* p = &tile[t * 65];
    float s = 0.f, l = 0.f, tracc = 0.f, gmax = 0.f;
    const bool longLane = (t & 16) != 0;
    const float A = longLane ? c.aL : c.aS, R = longLane ? c.rL : c.rS, omA = longLane ? c.omaL : c.omaS, omR = longLane ? c.omrL : c.omrS;
    long long t0 = clock64();
    for (int q = 0; q < n / 4; ++q) {
        float4 v = p[q & 63];
        float m[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = fabsf(m[j]);
            if (VARIANT == 0) {
                bool us = a > s; s = (us ? c.omaS : c.omrS) * a + (us ? c.aS : c.rS) * s;
                bool ul = a > l; l = (ul ? c.omaL : c.omrL) * a + (ul ? c.aL : c.rL) * l;
            } else if (VARIANT == 1) {
                unsigned us = gt_mask(a, s); s = bitsel(us, c.omaS, c.omrS) * a + bitsel(us, c.aS, c.rS) * s;
                unsigned ul = gt_mask(a, l); l = bitsel(ul, c.omaL, c.omrL) * a + bitsel(ul, c.aL, c.rL) * l;
            } else if (VARIANT == 2) {
                s = fmaxf(c.omaS * a + c.aS * s, c.omrS * a + c.rS * s);
                l = fmaxf(c.omaL * a + c.aL * l, c.omrL * a + c.rL * l);
            } else if (VARIANT == 3) {
                unsigned us = gt_mask(a, s); s = bitsel(us, c.omaS * a + c.aS * s, c.omrS * a + c.rS * s);
                unsigned ul = gt_mask(a, l); l = bitsel(ul, c.omaL * a + c.aL * l, c.omrL * a + c.rL * l);
            } else {
                bool us = a > s; s = (us ? omA : omR) * a + (us ? A : R) * s;
            }
            if (VARIANT != 4) {
                float tr = fmaxf(0.f, s - l);
                tracc += tr;
                gmax = fmaxf(gmax, tr);
            }
        }
        if (VARIANT == 4) { // tr for the quad's last sample only (cost model: the tr pass lives elsewhere)
            float o = __shfl_xor_sync(0xffffffffu, s, 16);
            tracc += fmaxf(0.f, s - o);
        }
    }
    long long t1 = clock64();
    out[t] = s + l + tracc + gmax;
    if (t == 0) cyc[0] = t1 - t0;
}
template <int V>
void run(const char* name, const float* x, int n, C c, float* out, long long* cyc)
{
    for (int warps : { 1, 2, 4, 8 }) {
        size_t smem = (size_t) warps * 32 * 65 * 16;
        cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        long long hc = 0;
        for (int rep = 0; rep < 2; ++rep) {
            k<V><<<1, warps * 32, smem>>>(x, n, c, out, cyc);
            cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
        }
        printf("%-44s warps %d: %.1f cycles/sample\n", name, warps, (double) hc / n);
    }
}
int main()
{
    const int n = 65536;
    float* x; float* out; long long* cyc;
    cudaMalloc(&x, sizeof(float) * 32 * n); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    float* h = new float[32 * n];
    unsigned r = 1; for (int i = 0; i < 32 * n; ++i) { r = r * 1664525u + 1013904223u; h[i] = ((r >> 8) & 0xffff) / 65536.f - 0.5f; }
    cudaMemcpy(x, h, sizeof(float) * 32 * n, cudaMemcpyHostToDevice);
    C c { 0.993f, 0.9993f, 0.9996f, 0.99993f, 0.007f, 0.0007f, 0.0004f, 0.00007f };
    run<0>("0 predicate select of coefficients", x, n, c, out, cyc);
    run<1>("1 mask + lop3 select of coefficients", x, n, c, out, cyc);
    run<2>("2 both candidates + fmax", x, n, c, out, cyc);
    run<3>("3 both candidates + mask/lop3 select", x, n, c, out, cyc);
    run<4>("4 one envelope per lane, predicate select", x, n, c, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
