// Microbenchmark: issue cost of the packed fp32x2 ALU ops of sm_100 (add/mul/fma.f32x2 -> FADD2 / FMUL2 / FFMA2) against
// their scalar forms.  Per thread: 8 independent accumulator chains (so latency is hidden), N iterations; 8 warps per
// scheduler.  Prints cycles per warp-instruction per scheduler and the useful flop rate relative to scalar.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float a, float b)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) { // scalar FFMA, 16 per iteration
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("fma.rn.ftz.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        } else if (MODE == 1) { // packed FFMA2, 8 per iteration (same 16 flop-lanes)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                asm volatile("{ .reg .b64 v, ka, kb; mov.b64 v, {%0, %1}; mov.b64 ka, {%2, %2}; mov.b64 kb, {%3, %3};\n"
                             "fma.rn.ftz.f32x2 v, v, ka, kb; mov.b64 {%0, %1}, v; }"
                             : "+f"(x[i]), "+f"(x[i + 1]) : "f"(a), "f"(b));
            }
        } else if (MODE == 2) { // scalar FMUL + FADD pairs (what -fmad=false code looks like): 16 mul + 16 add
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                asm volatile("mul.rn.ftz.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));
                asm volatile("add.rn.ftz.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
            }
        } else { // packed FMUL2 + FADD2: 8 + 8
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                asm volatile("{ .reg .b64 v, ka, kb; mov.b64 v, {%0, %1}; mov.b64 ka, {%2, %2}; mov.b64 kb, {%3, %3};\n"
                             "mul.rn.ftz.f32x2 v, v, ka; add.rn.ftz.f32x2 v, v, kb; mov.b64 {%0, %1}, v; }"
                             : "+f"(x[i]), "+f"(x[i + 1]) : "f"(a), "f"(b));
            }
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int instrPerIter)
{
    const int blocks = 148 * 4, threads = 256, iters = 20000;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 100, 0.999f, 0.001f);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    int clockKHz = 0;
    cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, 0);
    const double warpsPerSched = (double) blocks * threads / 32.0 / (148.0 * 4.0);
    const double cyc = ms * 1e-3 * clockKHz * 1e3;
    const double cycPerInstr = cyc / (warpsPerSched * iters * instrPerIter);
    printf("%-28s %8.3f ms  %.3f cycles per warp-instruction per scheduler (%d instr / iter), %.1f G flop-lanes/ms\n", name, ms, cycPerInstr,
           instrPerIter, (double) blocks * threads * iters * 16.0 / ms * 1e-9);
    cudaFree(out);
}

int main()
{
    run<0>("scalar FFMA x16", 16);
    run<1>("packed FFMA2 x8", 8);
    run<2>("scalar FMUL+FADD x16", 32);
    run<3>("packed FMUL2+FADD2 x8", 16);
    return 0;
}
