// Microbenchmark: what the C library's own tanhf / powf (csrc/jb_libm.h) cost on sm_100a next to the MUFU-based routines,
// and the DFMA rate that bounds powf.  Per thread: `CHAINS` independent evaluations per iteration (latency hidden when
// CHAINS is large), W warps per SM sub-partition.  Prints cycles per warp-call per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -ftz=true -prec-div=true -prec-sqrt=true \
//        -I../../juicy-audio-plugins_b200/csrc -o bin/exact_math exact_math.cu && bin/exact_math
#include <cstdio>
#include <cuda_runtime.h>
#include "jb_device.cuh"
#include "jb_libm.h"

using namespace jbdev;

template <int MODE, int CHAINS>
__global__ void k(float* out, int iters, float seed, float e)
{
    float x[CHAINS];
    double d[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        x[i] = seed + 0.013f * (float) ((threadIdx.x * 7 + i * 3) % 61);
        d[i] = (double) x[i];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) x[i] = tanh_fast(x[i]) + 0.37f;                       // MUFU tanh
            if (MODE == 1) x[i] = jblibm::tanhf_fdlibm(x[i]) + 0.37f;            // fdlibm tanhf
            if (MODE == 2) x[i] = pow_unit(x[i] * 0.5f, e) + 0.11f;              // MUFU pow
            if (MODE == 3) x[i] = jblibm::powf_glibc_pos(x[i] * 0.5f, e) + 0.11f; // glibc powf (double inside)
            if (MODE == 4) d[i] = fma(d[i], 0.999, 0.001);                       // DFMA
            if (MODE == 5) x[i] = fmaf(x[i], 0.999f, 0.001f);                    // FFMA
            if (MODE == 6) x[i] = jblibm::fdiv(1.0f, x[i] + 1.5f);               // IEEE fp32 division (fast-path sequence)
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i)
        s += x[i] + (float) d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int CHAINS>
void run(const char* name, int warpsPerSched)
{
    const int blocks = 148, threads = 128 * warpsPerSched, iters = 4000;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE, CHAINS><<<blocks, threads>>>(out, 50, 0.05f, 0.69f);
    cudaEventRecord(e0);
    k<MODE, CHAINS><<<blocks, threads>>>(out, iters, 0.05f, 0.69f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int mhz = 0;
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * mhz * 1e3;
    const double callsPerSched = (double) iters * CHAINS * warpsPerSched;
    printf("%-28s chains %2d warps/sched %d : %7.2f cycles per warp-call per scheduler (%.3f ms)\n", name, CHAINS, warpsPerSched,
           cycles / callsPerSched, ms);
    cudaFree(out);
}

int main()
{
    run<5, 8>("FFMA", 4);
    run<4, 8>("DFMA", 4);
    run<4, 1>("DFMA latency", 1);
    run<5, 1>("FFMA latency", 1);
    run<6, 8>("fdiv (IEEE sequence)", 4);
    run<0, 4>("tanh_fast (MUFU)", 4);
    run<1, 4>("tanhf_fdlibm", 4);
    run<1, 1>("tanhf_fdlibm latency", 1);
    run<2, 4>("pow_unit (MUFU)", 4);
    run<3, 4>("powf_glibc_pos", 4);
    run<3, 1>("powf_glibc_pos latency", 1);
    run<3, 2>("powf_glibc_pos 2 chains", 2);
    return 0;
}
