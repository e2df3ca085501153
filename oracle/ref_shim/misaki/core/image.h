// Stand-in for include/misaki/core/image.h (the real one is written through OpenImageIO, image.cpp:20-43): a plain
// pixel-interleaved float buffer with the accessors HDRFilm::image uses (operator()(x, y, channel), clamped like the
// reference's).  write() is not part of the pinned build.  TEST INFRASTRUCTURE.
#pragma once
#include "msk_ref_prelude.h"
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>
namespace misaki {
class Image {
public:
    Image(const Eigen::Vector2i &size, const std::vector<std::string> channels, uint8_t * = nullptr)
        : m_size(size), m_names(channels), m_pixels((size_t) size.x() * size.y() * channels.size(), 0.f) {}
    float &operator()(int x, int y, int ch) { return m_pixels[index(x, y, ch)]; }
    const float &operator()(int x, int y, int ch) const { return m_pixels[index(x, y, ch)]; }
    void write(const fs::path &) { Throw("Image::write is outside the pinned build"); }
    const std::vector<std::string> &channel_names() const { return m_names; }
    const std::vector<float> &pixels() const { return m_pixels; } // H x W x channels
private:
    size_t index(int x, int y, int ch) const {
        x = std::clamp(x, 0, m_size.x() - 1);
        y = std::clamp(y, 0, m_size.y() - 1);
        return ((size_t) y * m_size.x() + x) * m_names.size() + ch;
    }
    Eigen::Vector2i m_size;
    std::vector<std::string> m_names;
    std::vector<float> m_pixels;
};
} // namespace misaki
