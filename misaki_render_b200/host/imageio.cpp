// Film development and image output (the step after the hot path).
//   develop_xyzaw   HDRFilm::image, reference src/librender/films/hdrfilm.cpp:48-90 + xyz_to_srgb spectrum.h:138-143
//   write_exr_rgba  what Image::write does through OpenImageIO (image.cpp:20-43): an RGBA float32 OpenEXR file;
//                   written here directly as an uncompressed scanline EXR (OpenImageIO is not available)
//   write_pfm_rgb   the reference's "pfm" file_format
#include "render.h"

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

void write_exr_rgba(const std::string &filename, const float *rgba, uint32_t width, uint32_t height) {
    Buf b;
    b.i32(20000630); // magic
    b.i32(2);        // version 2, single-part scanline
    { // channels, alphabetical: A B G R, FLOAT
        Buf c;
        for (const char *name : { "A", "B", "G", "R" }) { c.str(name); c.i32(2); c.u8(0); c.u8(0); c.u8(0); c.u8(0); c.i32(1); c.i32(1); }
        c.u8(0);
        b.attr("channels", "chlist", c.d.data(), (int32_t) c.d.size());
    }
    uint8_t comp = 0; b.attr("compression", "compression", &comp, 1);
    int32_t window[4] = { 0, 0, (int32_t) width - 1, (int32_t) height - 1 };
    b.attr("dataWindow", "box2i", window, 16);
    b.attr("displayWindow", "box2i", window, 16);
    uint8_t order = 0; b.attr("lineOrder", "lineOrder", &order, 1);
    float par = 1.f; b.attr("pixelAspectRatio", "float", &par, 4);
    float center[2] = { 0.f, 0.f }; b.attr("screenWindowCenter", "v2f", center, 8);
    float sww = 1.f; b.attr("screenWindowWidth", "float", &sww, 4);
    b.u8(0); // end of header
    const size_t row_bytes = (size_t) width * 4 * 4, table_pos = b.d.size();
    uint64_t offset = table_pos + (uint64_t) height * 8;
    for (uint32_t y = 0; y < height; ++y) { b.u64(offset); offset += 8 + row_bytes; }
    std::vector<float> row((size_t) width * 4);
    static const int src_of[4] = { 3, 2, 1, 0 }; // A B G R <- rgba
    for (uint32_t y = 0; y < height; ++y) {
        b.i32((int32_t) y);
        b.i32((int32_t) row_bytes);
        for (int c = 0; c < 4; ++c)
            for (uint32_t x = 0; x < width; ++x) row[(size_t) c * width + x] = rgba[((size_t) y * width + x) * 4 + src_of[c]];
        b.bytes(row.data(), row_bytes);
    }
    FILE *f = fopen(filename.c_str(), "wb");
    if (!f) Throw("Could not open \"%s\" for writing", filename.c_str());
    bool ok = fwrite(b.d.data(), 1, b.d.size(), f) == b.d.size();
    fclose(f);
    if (!ok) Throw("Error while writing \"%s\"", filename.c_str());
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
