// Film development and image output (the step after the hot path).
//   develop_xyzaw   HDRFilm::image, reference src/librender/films/hdrfilm.cpp:48-90 + xyz_to_srgb spectrum.h:138-143
//   write_exr_rgba  what Image::write does through OpenImageIO (image.cpp:20-43): an RGBA float32 OpenEXR file;
//                   written here directly as an uncompressed scanline EXR (OpenImageIO is not available)
//   write_pfm_rgb   the reference's "pfm" file_format
#include "render.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace misaki {

void develop_xyzaw(const float *film, size_t n, float *rgba) {
    for (size_t i = 0; i < n; ++i) {
        const float *p = film + i * 5;
        float r = 3.240479f * p[0] + -1.537150f * p[1] + -0.498535f * p[2];
        float g = -0.969256f * p[0] + 1.875991f * p[1] + 0.041556f * p[2];
        float b = 0.055648f * p[0] + -0.204043f * p[1] + 1.057311f * p[2];
        float inv = p[4] != 0.f ? 1.f / p[4] : 0.f;
        rgba[i * 4 + 0] = r * inv; rgba[i * 4 + 1] = g * inv; rgba[i * 4 + 2] = b * inv; rgba[i * 4 + 3] = p[3] * inv;
    }
}

// HDRFilm::image with AOV channels (hdrfilm.cpp:50-88): RGBA as above, then every channel past W divided by W.
// film: n x nch floats, out: n x (nch - 1) floats.
void develop_channels(const float *film, size_t n, size_t nch, float *out) {
    for (size_t i = 0; i < n; ++i) {
        const float *p = film + i * nch;
        float *o = out + i * (nch - 1);
        float r = 3.240479f * p[0] + -1.537150f * p[1] + -0.498535f * p[2];
        float g = -0.969256f * p[0] + 1.875991f * p[1] + 0.041556f * p[2];
        float b = 0.055648f * p[0] + -0.204043f * p[1] + 1.057311f * p[2];
        float inv = p[4] != 0.f ? 1.f / p[4] : 0.f;
        o[0] = r * inv; o[1] = g * inv; o[2] = b * inv; o[3] = p[3] * inv;
        for (size_t ch = 5; ch < nch; ++ch) o[ch - 1] = p[ch] * inv;
    }
}

namespace {
struct Buf {
    std::vector<uint8_t> d;
    void bytes(const void *p, size_t n) { const uint8_t *b = (const uint8_t *) p; d.insert(d.end(), b, b + n); }
    void str(const char *s) { bytes(s, strlen(s) + 1); }
    void u8(uint8_t v) { d.push_back(v); }
    void i32(int32_t v) { bytes(&v, 4); }
    void u64(uint64_t v) { bytes(&v, 8); }
    void f32(float v) { bytes(&v, 4); }
    void attr(const char *name, const char *type, const void *data, int32_t size) { str(name); str(type); i32(size); bytes(data, size); }
};
} // namespace

// `pixels`: height x width x names.size() floats, pixel-interleaved in the order of `names`.  OpenEXR stores the
// channels of a scanline one after the other in alphabetical order of their names.
void write_exr_channels(const std::string &filename, const std::vector<std::string> &names, const float *pixels, uint32_t width,
                        uint32_t height) {
    const size_t nch = names.size();
    std::vector<size_t> order(nch);
    for (size_t i = 0; i < nch; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return names[a] < names[b]; });
    for (size_t i = 1; i < nch; ++i)
        if (names[order[i]] == names[order[i - 1]]) Throw("write_exr: duplicate channel name \"%s\"", names[order[i]].c_str());
    Buf b;
    b.i32(20000630); // magic
    b.i32(2);        // version 2, single-part scanline
    { // channel list, FLOAT
        Buf c;
        for (size_t i = 0; i < nch; ++i) {
            const std::string &name = names[order[i]];
            if (name.empty() || name.size() > 255) Throw("write_exr: invalid channel name");
            c.str(name.c_str()); c.i32(2); c.u8(0); c.u8(0); c.u8(0); c.u8(0); c.i32(1); c.i32(1);
        }
        c.u8(0);
        b.attr("channels", "chlist", c.d.data(), (int32_t) c.d.size());
    }
    uint8_t comp = 0; b.attr("compression", "compression", &comp, 1);
    int32_t window[4] = { 0, 0, (int32_t) width - 1, (int32_t) height - 1 };
    b.attr("dataWindow", "box2i", window, 16);
    b.attr("displayWindow", "box2i", window, 16);
    uint8_t line_order = 0; b.attr("lineOrder", "lineOrder", &line_order, 1);
    float par = 1.f; b.attr("pixelAspectRatio", "float", &par, 4);
    float center[2] = { 0.f, 0.f }; b.attr("screenWindowCenter", "v2f", center, 8);
    float sww = 1.f; b.attr("screenWindowWidth", "float", &sww, 4);
    b.u8(0); // end of header
    const size_t row_bytes = (size_t) width * nch * 4, table_pos = b.d.size();
    uint64_t offset = table_pos + (uint64_t) height * 8;
    for (uint32_t y = 0; y < height; ++y) { b.u64(offset); offset += 8 + row_bytes; }
    std::vector<float> row((size_t) width * nch);
    for (uint32_t y = 0; y < height; ++y) {
        b.i32((int32_t) y);
        b.i32((int32_t) row_bytes);
        for (size_t c = 0; c < nch; ++c)
            for (uint32_t x = 0; x < width; ++x) row[c * width + x] = pixels[((size_t) y * width + x) * nch + order[c]];
        b.bytes(row.data(), row_bytes);
    }
    FILE *f = fopen(filename.c_str(), "wb");
    if (!f) Throw("Could not open \"%s\" for writing", filename.c_str());
    bool ok = fwrite(b.d.data(), 1, b.d.size(), f) == b.d.size();
    fclose(f);
    if (!ok) Throw("Error while writing \"%s\"", filename.c_str());
}

void write_exr_rgba(const std::string &filename, const float *rgba, uint32_t width, uint32_t height) {
    write_exr_channels(filename, { "R", "G", "B", "A" }, rgba, width, height);
}

void write_pfm_rgb(const std::string &filename, const float *rgba, uint32_t width, uint32_t height) {
    FILE *f = fopen(filename.c_str(), "wb");
    if (!f) Throw("Could not open \"%s\" for writing", filename.c_str());
    fprintf(f, "PF\n%u %u\n-1.0\n", width, height);
    std::vector<float> row((size_t) width * 3);
    for (uint32_t yy = 0; yy < height; ++yy) { // PFM stores the bottom scanline first
        uint32_t y = height - 1 - yy;
        for (uint32_t x = 0; x < width; ++x)
            for (int c = 0; c < 3; ++c) row[(size_t) x * 3 + c] = rgba[((size_t) y * width + x) * 4 + c];
        fwrite(row.data(), sizeof(float), row.size(), f);
    }
    fclose(f);
}

} // namespace misaki
