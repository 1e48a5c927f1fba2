// ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false (nvcc and -Xptxas):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -cubin -o fz.cubin f32x2_fuse.cu && cuobjdump -sass fz.cubin | grep "FMUL2\|FADD2\|FFMA2"
// prints one FMUL2 and one FFMA2, no FADD2.
struct F2 { float x, y; };
__device__ __forceinline__ F2 mul2(F2 a, F2 b)
{
    F2 r;
    asm("{ .reg .b64 pa, pb, pc; mov.b64 pa, {%2, %3}; mov.b64 pb, {%4, %5}; mul.rn.ftz.f32x2 pc, pa, pb; mov.b64 {%0, %1}, pc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b)
{
    F2 r;
    asm("{ .reg .b64 pa, pb, pc; mov.b64 pa, {%2, %3}; mov.b64 pb, {%4, %5}; add.rn.ftz.f32x2 pc, pa, pb; mov.b64 {%0, %1}, pc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__global__ void k(float* o, float a, float b, float c, float d)
{
    F2 x { o[0], o[1] }, y { o[2], o[3] };
    F2 r = add2(mul2(F2 { a, b }, x), mul2(F2 { c, d }, y));
    o[4] = r.x; o[5] = r.y;
}
