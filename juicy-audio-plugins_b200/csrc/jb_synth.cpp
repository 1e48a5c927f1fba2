// jb_synth.cpp -- host generator of the seeded synthetic clips:
// exponential sine sweep, white noise, impulse train, drum hit.  Same formulas as
// jb_synth_kernel in jb_kernels.cu; used by tests and the CPU-baseline legs so the
// oracle and the engine see identical input bits.
#include "jb_params.h"

#include <cmath>
#include <cstdint>

namespace jb {
namespace {

constexpr float kPi = 3.14159265358979323846f;

inline uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// uniform in [-1, 1), exactly representable
inline float unitNoise(uint32_t seed, uint32_t ch, uint32_t n)
{
    const uint32_t h = hash32(seed ^ hash32(n * 2u + ch + 0x9E3779B9u));
    return (float) (h >> 8) * (1.0f / 8388608.0f) - 1.0f;
}

} // namespace

void synthFillHost(float* audio, int kind, long long firstClip, int nClips, int nCh, int nSamples,
                   double sampleRate, unsigned int baseSeed)
{
    const float sr = (float) sampleRate;
    for (int clipLocal = 0; clipLocal < nClips; ++clipLocal) {
        const long long clipId = firstClip + clipLocal;
        const uint32_t seed = baseSeed ^ ((uint32_t) clipId * 0x9E3779B9u);
        const int k = kind == 4 ? (int) (clipId & 3) : kind;
        float* left = audio + (long long) clipLocal * nCh * nSamples;
        float* right = nCh > 1 ? left + nSamples : nullptr;
        const uint32_t h = hash32(seed);
        const int onset = 480 + (int) (h % 4800u);
        const float f0 = 45.0f + 45.0f * (float) ((h >> 13) & 1023u) / 1023.0f;
        const int period = 2400 + 37 * (int) (clipId & 63);
        for (int n = 0; n < nSamples; ++n) {
            float l = 0.0f, r = 0.0f;
            if (k == 0) {
                const float T = (float) nSamples / sr;
                const float K = std::log(1000.0f);
                const float t = (float) n / sr;
                const float ph = 2.0f * kPi * 20.0f * T / K * (std::exp(t / T * K) - 1.0f);
                l = 0.5f * std::sin(ph);
                r = 0.5f * std::sin(ph + 0.3f);
            } else if (k == 1) {
                l = 0.25f * unitNoise(seed, 0u, (uint32_t) n);
                r = 0.5f * (l + 0.25f * unitNoise(seed, 1u, (uint32_t) n));
            } else if (k == 2) {
                l = (n % period) == 0 ? 0.9f : 0.0f;
                r = (n >= 7 && ((n - 7) % period) == 0) ? 0.9f : 0.0f;
            } else if (n >= onset) {
                const float m = (float) (n - onset);
                const float body = 0.8f * std::exp(-m / 2400.0f) * std::sin(2.0f * kPi * f0 * m / sr);
                const float burst = 0.4f * std::exp(-m / 600.0f);
                l = body + burst * unitNoise(seed, 0u, (uint32_t) n);
                r = 0.8f * l + 0.2f * burst * unitNoise(seed, 1u, (uint32_t) n);
            }
            left[n] = l;
            if (right)
                right[n] = r;
        }
    }
}

} // namespace jb
