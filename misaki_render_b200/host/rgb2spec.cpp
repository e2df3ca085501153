// RGB -> spectral-coefficient lookup (setup time, once per texture).
// Restates rgb2spec_load / rgb2spec_find_interval / rgb2spec_fetch of the reference's vendored
// ext/rgb2spec/rgb2spec.c:12-47,59-119 (Jakob & Hanika 2019) as used by src/librender/srgb.cpp:11-30.
// The table is the file the reference generates at build time with `rgb2spec_opt 64`
// (ext/rgb2spec/CMakeLists.txt:41-46) and resolves as "data/srgb.coeff".
#include "render.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace misaki {
namespace {

struct Model {
    uint32_t res = 0;
    std::vector<float> scale, data;
};

Model *load_model(const std::string &filename) {
    FILE *f = fopen(filename.c_str(), "rb");
    if (!f) return nullptr;
    char header[4];
    Model *m = new Model;
    bool ok = fread(header, 4, 1, f) == 1 && memcmp(header, "SPEC", 4) == 0 && fread(&m->res, sizeof(uint32_t), 1, f) == 1 &&
              m->res >= 2 && m->res <= 4096;
    if (ok) {
        size_t size_scale = m->res, size_data = (size_t) m->res * m->res * m->res * 3 * 3;
        m->scale.resize(size_scale);
        m->data.resize(size_data);
        ok = fread(m->scale.data(), sizeof(float), size_scale, f) == size_scale && fread(m->data.data(), sizeof(float), size_data, f) == size_data;
    }
    fclose(f);
    if (!ok) { delete m; return nullptr; }
    return m;
}

int find_interval(const float *values, int size_, float x) { // rgb2spec.c:59-75
    int left = 0, last_interval = size_ - 2, size = last_interval;
    while (size > 0) {
        int half = size >> 1, middle = left + half + 1;
        if (values[middle] <= x) { left = middle; size -= half + 1; }
        else size = half;
    }
    return left < last_interval ? left : last_interval;
}

Model *g_model = nullptr;
std::mutex g_mutex;

} // namespace

Color3 srgb_model_fetch(const Color3 &c) {
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        if (!g_model) {
            std::string fname = get_file_resolver()->resolve("data/srgb.coeff");
            Log(Info, "Loading spectral upsampling model \"data/srgb.coeff\" .. ");
            g_model = load_model(fname);
            if (!g_model) Throw("Could not load sRGB-to-spectrum upsampling model ('data/srgb.coeff')");
        }
    }
    const Model &model = *g_model;
    int i = 0, res = (int) model.res;
    float rgb_[3] = { c.r, c.g, c.b }, rgb[3];
    for (int j = 0; j < 3; ++j) rgb[j] = std::fmax(std::fmin(rgb_[j], 1.f), 0.f);
    for (int j = 1; j < 3; ++j)
        if (rgb[j] >= rgb[i]) i = j;
    float z = rgb[i], scale = (res - 1) / z, x = rgb[(i + 1) % 3] * scale, y = rgb[(i + 2) % 3] * scale;
    // (uint32_t) of NaN (black: 0 * inf) is undefined in C; x86 yields 0 for the cases that occur
    auto to_u32 = [](float v) -> uint32_t { return (std::isfinite(v) && v > 0.f) ? (uint32_t) v : 0u; };
    uint32_t xi = std::min(to_u32(x), (uint32_t) (res - 2)), yi = std::min(to_u32(y), (uint32_t) (res - 2)),
             zi = (uint32_t) find_interval(model.scale.data(), res, z);
    size_t offset = ((((size_t) i * res + zi) * res + yi) * res + xi) * 3, dx = 3, dy = 3 * (size_t) res, dz = 3 * (size_t) res * res;
    float x1 = x - xi, x0 = 1.f - x1, y1 = y - yi, y0 = 1.f - y1,
          z1 = (z - model.scale[zi]) / (model.scale[zi + 1] - model.scale[zi]), z0 = 1.f - z1;
    float out[3];
    const float *d = model.data.data();
    for (int j = 0; j < 3; ++j) {
        out[j] = ((d[offset] * x0 + d[offset + dx] * x1) * y0 + (d[offset + dy] * x0 + d[offset + dy + dx] * x1) * y1) * z0 +
                 ((d[offset + dz] * x0 + d[offset + dz + dx] * x1) * y0 + (d[offset + dz + dy] * x0 + d[offset + dz + dy + dx] * x1) * y1) * z1;
        offset++;
    }
    return Color3{ out[0], out[1], out[2] };
}

} // namespace misaki
