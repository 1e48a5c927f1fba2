// jb_wav.cpp -- RIFF/WAVE files <-> the planar fp32 [channel][sample] buffers the engine renders.
//
// The reference has no file I/O of its own: the DAW hands its plugins decoded fp32 AudioBuffers.  An offline batch
// renderer has to do that step itself, so the library reads and writes the one format every DAW exports: PCM 16 / 24 /
// 32-bit integer and 32-bit IEEE float, plain or WAVE_FORMAT_EXTENSIBLE headers, any chunk order, odd chunks padded.
// Sample conversion follows what JUCE's WavAudioFormat / AudioData converters do, so that a file decoded here holds the
// floats the plugin would have seen inside a JUCE host: integer -> float is value / 2^(bits-1); float -> integer is
// round(value * 2^(bits-1)) limited to +-(2^(bits-1) - 1).  Host code only; no GPU involved.
#include "../../include/juicy_batch.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_wavError;

int wavFail(const std::string& msg)
{
    g_wavError = msg;
    return JB_ERR_ARG;
}

uint32_t rd32(const unsigned char* p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16) | ((uint32_t) p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t) (p[0] | (p[1] << 8)); }
void wr32(unsigned char* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (unsigned char) (v >> (8 * i)); }
void wr16(unsigned char* p, uint16_t v) { p[0] = (unsigned char) v; p[1] = (unsigned char) (v >> 8); }

struct Layout {
    int channels = 0, bits = 0, blockAlign = 0;
    bool isFloat = false;
    double rate = 0.0;
    long dataOffset = 0;
    uint32_t dataBytes = 0;
};

// Walks the chunk list of an open file; leaves the layout of "fmt " and the position of "data".
int parse(FILE* f, Layout* out)
{
    unsigned char hdr[12];
    if (std::fread(hdr, 1, 12, f) != 12 || std::memcmp(hdr, "RIFF", 4) != 0 || std::memcmp(hdr + 8, "WAVE", 4) != 0)
        return wavFail("not a RIFF/WAVE file");
    bool haveFmt = false;
    for (;;) {
        unsigned char ch[8];
        if (std::fread(ch, 1, 8, f) != 8)
            break;
        const uint32_t size = rd32(ch + 4);
        const long body = std::ftell(f);
        if (std::memcmp(ch, "fmt ", 4) == 0) {
            unsigned char fm[40] = {};
            const size_t want = size < 40 ? size : 40;
            if (size < 16 || std::fread(fm, 1, want, f) != want)
                return wavFail("truncated fmt chunk");
            uint16_t tag = rd16(fm);
            out->channels = rd16(fm + 2);
            out->rate = (double) rd32(fm + 4);
            out->blockAlign = rd16(fm + 12);
            out->bits = rd16(fm + 14);
            if (tag == 0xFFFE && size >= 40) // WAVE_FORMAT_EXTENSIBLE: the sub-format GUID starts with the real tag
                tag = rd16(fm + 24);
            if (tag != 1 && tag != 3)
                return wavFail("unsupported WAVE format tag (PCM and IEEE float only)");
            out->isFloat = tag == 3;
            haveFmt = true;
        } else if (std::memcmp(ch, "data", 4) == 0) {
            out->dataOffset = body;
            out->dataBytes = size;
            if (!haveFmt)
                return wavFail("data chunk before fmt chunk");
            return JB_OK;
        }
        if (std::fseek(f, body + (long) size + (long) (size & 1u), SEEK_SET) != 0)
            break;
    }
    return wavFail("no data chunk");
}

bool supported(const Layout& l)
{
    if (l.channels < 1 || l.blockAlign != l.channels * (l.bits / 8))
        return false;
    return l.isFloat ? l.bits == 32 : (l.bits == 16 || l.bits == 24 || l.bits == 32);
}

} // namespace

extern "C" {

const char* jb_wav_last_error(void) { return g_wavError.c_str(); }

int jb_wav_info_read(const char* path, jb_wav_info* out)
{
    if (path == nullptr || out == nullptr)
        return wavFail("null argument");
    FILE* f = std::fopen(path, "rb");
    if (f == nullptr)
        return wavFail(std::string("cannot open ") + path);
    Layout l;
    const int rc = parse(f, &l);
    std::fclose(f);
    if (rc != JB_OK)
        return rc;
    if (!supported(l))
        return wavFail("unsupported sample layout");
    out->n_channels = l.channels;
    out->n_samples = (int) (l.dataBytes / (uint32_t) l.blockAlign);
    out->sample_rate = l.rate;
    out->bits_per_sample = l.bits;
    out->is_float = l.isFloat ? 1 : 0;
    return JB_OK;
}

int jb_wav_read(const char* path, float* h_planar, int n_channels, int n_samples)
{
    if (path == nullptr || h_planar == nullptr)
        return wavFail("null argument");
    FILE* f = std::fopen(path, "rb");
    if (f == nullptr)
        return wavFail(std::string("cannot open ") + path);
    Layout l;
    int rc = parse(f, &l);
    if (rc == JB_OK && !supported(l))
        rc = wavFail("unsupported sample layout");
    if (rc == JB_OK && (l.channels != n_channels || (int) (l.dataBytes / (uint32_t) l.blockAlign) < n_samples || n_samples < 0))
        rc = wavFail("file shape does not match the requested channels / samples");
    if (rc != JB_OK) {
        std::fclose(f);
        return rc;
    }
    std::fseek(f, l.dataOffset, SEEK_SET);
    const int bytes = l.bits / 8;
    std::vector<unsigned char> frames((size_t) 4096 * (size_t) l.blockAlign);
    int done = 0;
    while (done < n_samples) {
        const int take = n_samples - done < 4096 ? n_samples - done : 4096;
        if (std::fread(frames.data(), (size_t) l.blockAlign, (size_t) take, f) != (size_t) take) {
            std::fclose(f);
            return wavFail("truncated data chunk");
        }
        for (int i = 0; i < take; ++i)
            for (int c = 0; c < n_channels; ++c) {
                const unsigned char* p = frames.data() + (size_t) i * l.blockAlign + (size_t) c * bytes;
                float v;
                if (l.isFloat) {
                    const uint32_t u = rd32(p);
                    std::memcpy(&v, &u, 4);
                } else if (l.bits == 16) {
                    v = (float) (int16_t) rd16(p) * (1.0f / 32768.0f);
                } else if (l.bits == 24) {
                    const int32_t s = (int32_t) ((uint32_t) p[0] << 8 | (uint32_t) p[1] << 16 | (uint32_t) p[2] << 24) >> 8;
                    v = (float) s * (1.0f / 8388608.0f);
                } else {
                    v = (float) ((double) (int32_t) rd32(p) * (1.0 / 2147483648.0));
                }
                h_planar[(size_t) c * (size_t) n_samples + (size_t) (done + i)] = v;
            }
        done += take;
    }
    std::fclose(f);
    return JB_OK;
}

int jb_wav_write(const char* path, const float* h_planar, int n_channels, int n_samples, double sample_rate, int bits_per_sample,
                 int is_float)
{
    if (path == nullptr || h_planar == nullptr || n_channels < 1 || n_samples < 0 || sample_rate <= 0.0)
        return wavFail("bad argument");
    const bool flt = is_float != 0;
    if (flt ? bits_per_sample != 32 : (bits_per_sample != 16 && bits_per_sample != 24 && bits_per_sample != 32))
        return wavFail("bits_per_sample: 16 / 24 / 32 integer or 32 float");
    const int bytes = bits_per_sample / 8, align = bytes * n_channels;
    const uint64_t dataBytes = (uint64_t) align * (uint64_t) n_samples;
    if (dataBytes > 0xFFFFFF00ull)
        return wavFail("too long for a RIFF file");
    FILE* f = std::fopen(path, "wb");
    if (f == nullptr)
        return wavFail(std::string("cannot create ") + path);
    unsigned char h[44];
    std::memcpy(h, "RIFF", 4);
    wr32(h + 4, (uint32_t) (36 + dataBytes + (dataBytes & 1u)));
    std::memcpy(h + 8, "WAVEfmt ", 8);
    wr32(h + 16, 16);
    wr16(h + 20, flt ? 3 : 1);
    wr16(h + 22, (uint16_t) n_channels);
    wr32(h + 24, (uint32_t) std::llround(sample_rate));
    wr32(h + 28, (uint32_t) std::llround(sample_rate) * (uint32_t) align);
    wr16(h + 32, (uint16_t) align);
    wr16(h + 34, (uint16_t) bits_per_sample);
    std::memcpy(h + 36, "data", 4);
    wr32(h + 40, (uint32_t) dataBytes);
    bool ok = std::fwrite(h, 1, 44, f) == 44;
    std::vector<unsigned char> frames((size_t) 4096 * (size_t) align);
    const double scale = std::ldexp(1.0, bits_per_sample - 1), top = scale - 1.0;
    for (int done = 0; ok && done < n_samples;) {
        const int take = n_samples - done < 4096 ? n_samples - done : 4096;
        for (int i = 0; i < take; ++i)
            for (int c = 0; c < n_channels; ++c) {
                const float v = h_planar[(size_t) c * (size_t) n_samples + (size_t) (done + i)];
                unsigned char* p = frames.data() + (size_t) i * align + (size_t) c * bytes;
                if (flt) {
                    uint32_t u;
                    std::memcpy(&u, &v, 4);
                    wr32(p, u);
                } else {
                    double q = std::nearbyint((double) v * scale);
                    q = q > top ? top : (q < -top ? -top : q);
                    const int32_t s = (int32_t) q;
                    for (int b = 0; b < bytes; ++b)
                        p[b] = (unsigned char) ((uint32_t) s >> (8 * b));
                }
            }
        ok = std::fwrite(frames.data(), (size_t) align, (size_t) take, f) == (size_t) take;
        done += take;
    }
    if (ok && (dataBytes & 1u)) {
        const unsigned char pad = 0;
        ok = std::fwrite(&pad, 1, 1, f) == 1;
    }
    ok = std::fclose(f) == 0 && ok;
    return ok ? JB_OK : wavFail(std::string("write failed: ") + path);
}

} // extern "C"
