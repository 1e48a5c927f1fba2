// jb_single.cuh -- one plugin per launch: kernel template and launcher shared by the jb_single_*.cu translation units
// (several, so that nvcc compiles the plugin kernels in parallel).
#pragma once
#include "jb_lane.cuh"

namespace {

// One plugin per launch (single-plugin engines, and every launch of a chain rendered plugin by plugin): the same two
// sweeps per block with the slot index a compile-time 0, one kernel per plugin kind.  What that buys over the generic
// kernel above: every coefficient is a constant-bank operand (the generic kernel re-read them through a pointer for each
// sample, `LD.E` -- profiles/r01_s6_tex_lane_before.txt), the code of one plugin is a few thousand instructions instead of
// 315 k (instruction-cache misses were 18 % of Texture-metal's stall samples), and the register budget is per plugin:
// MIN_CTAS = 8 gives the heavy plugins 255 registers when the batch is at most 8 warps per SM anyway.
template <class Main, class Pre, int MIN_CTAS>
__global__ void __launch_bounds__(JB_CTA_THREADS, MIN_CTAS) jb_single_kernel(const __grid_constant__ ProcArgs a)
{
    const long long lane = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= a.nClips)
        return;
    const long long clip = a.clipMap != nullptr ? (long long) a.clipMap[lane] : lane;
    int blockAbs = a.histFirstBlock;
    unsigned tmaCount = 0; // TMA stages loaded so far by this warp (octets == 3)
    if (!Main::kHeavy && a.octets == 3)
        tma_tiles_init();
    for (int pos = 0; pos < a.nSamples; pos += a.blockSize, ++blockAbs) {
        const int n = min(a.blockSize, a.nSamples - pos);
        if constexpr (std::is_same<Main, MainInfer>::value) {
            // trim = 0 dB: applyGain(1) leaves the buffer as it is, so both analyze() calls of the block see the same
            // samples and share their state-independent sums (a warp-uniform choice: the gain mode is a launch constant)
            if (a.slot[0].c.infer.gainMode == 0) {
                BlockStats sums;
                sweep<MainNone, Pre>(a, clip, -1, pos, n, blockAbs, &sums, &tmaCount);
                sweep<Main, PreNone, false, true>(a, clip, 0, pos, n, blockAbs, &sums, &tmaCount);
                continue;
            }
        }
        sweep<MainNone, Pre>(a, clip, -1, pos, n, blockAbs, nullptr, &tmaCount);
        sweep<Main, PreNone>(a, clip, 0, pos, n, blockAbs, nullptr, &tmaCount);
    }
}

template <class Main, class Pre>
cudaError_t launch_single(const ProcArgs& a, int grid, cudaStream_t stream)
{
    // heavy plugins always take the four-samples-per-trip path, whatever `octets` says: size the ring for it
    const size_t smem = lane_smem_bytes(Main::kHeavy ? 0 : a.octets);
    // warps per SM this batch can supply; the 255-register variant holds at most 8
    if (Main::kHeavy && grid <= 8 * 148)
        jb_single_kernel<Main, Pre, 8><<<grid, JB_CTA_THREADS, smem, stream>>>(a);
    else
        jb_single_kernel<Main, Pre, 16><<<grid, JB_CTA_THREADS, smem, stream>>>(a);
    return cudaGetLastError();
}

} // namespace
