// bmp_io.h — minimal uncompressed 24-bit BMP / raw plane I/O for the headless drivers.
// Replaces what the reference's drivers get from OpenCV 2.x (imread / VideoCapture / imshow:
// image_io.cpp:95-96, video_io.cpp:76-86), which this image does not have.  Pixels are returned as
// OpenCV's Mat.data would hold them: top-down rows, tightly packed interleaved BGR.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

namespace bmpio {

inline bool read_bmp(const std::string &path, std::vector<uint8_t> &bgr, int &rows, int &cols)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint8_t h[54];
    if (fread(h, 1, 54, f) != 54 || h[0] != 'B' || h[1] != 'M') { fclose(f); return false; }
    auto u32 = [&](int o) { return (uint32_t)h[o] | (uint32_t)h[o + 1] << 8 | (uint32_t)h[o + 2] << 16 | (uint32_t)h[o + 3] << 24; };
    const uint32_t off = u32(10);
    const int32_t w = (int32_t)u32(18), hh = (int32_t)u32(22);
    const int bpp = h[28] | h[29] << 8, comp = (int)u32(30);
    if ((bpp != 24 && bpp != 32) || comp != 0 || w <= 0 || hh == 0) { fclose(f); return false; }
    cols = w;
    rows = hh < 0 ? -hh : hh;
    const size_t bypp = bpp / 8, pitch = ((size_t)w * bypp + 3) & ~(size_t)3;
    std::vector<uint8_t> line(pitch);
    bgr.assign((size_t)rows * cols * 3, 0);
    fseek(f, off, SEEK_SET);
    for (int r = 0; r < rows; ++r) {
        if (fread(line.data(), 1, pitch, f) != pitch) { fclose(f); return false; }
        const int y = hh < 0 ? r : rows - 1 - r;  // positive height = bottom-up
        for (int x = 0; x < cols; ++x) memcpy(&bgr[((size_t)y * cols + x) * 3], &line[x * bypp], 3);
    }
    fclose(f);
    return true;
}

inline bool write_bmp(const std::string &path, const uint8_t *bgr, int rows, int cols)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const size_t pitch = ((size_t)cols * 3 + 3) & ~(size_t)3, img = pitch * rows;
    uint8_t h[54] = {0};
    auto put = [&](int o, uint32_t v) { h[o] = v; h[o + 1] = v >> 8; h[o + 2] = v >> 16; h[o + 3] = v >> 24; };
    h[0] = 'B'; h[1] = 'M';
    put(2, (uint32_t)(54 + img)); put(10, 54); put(14, 40); put(18, (uint32_t)cols); put(22, (uint32_t)rows);
    h[26] = 1; h[28] = 24; put(34, (uint32_t)img);
    fwrite(h, 1, 54, f);
    std::vector<uint8_t> line(pitch, 0);
    for (int r = rows - 1; r >= 0; --r) {
        memcpy(line.data(), bgr + (size_t)r * cols * 3, (size_t)cols * 3);
        fwrite(line.data(), 1, pitch, f);
    }
    fclose(f);
    return true;
}

inline bool write_raw(const std::string &path, const void *data, size_t bytes)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(data, 1, bytes, f) == bytes;
    fclose(f);
    return ok;
}

// float disparity plane -> 8-bit grey BMP, min..max stretched (what the reference's imshow of a
// normalised Mat displays, image_io.cpp:333-336)
inline bool write_plane_bmp(const std::string &path, const float *p, int rows, int cols)
{
    float lo = p[0], hi = p[0];
    for (size_t i = 0; i < (size_t)rows * cols; ++i) { lo = p[i] < lo ? p[i] : lo; hi = p[i] > hi ? p[i] : hi; }
    const float sc = hi > lo ? 255.0f / (hi - lo) : 0.0f;
    std::vector<uint8_t> g((size_t)rows * cols * 3);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) g[3 * i] = g[3 * i + 1] = g[3 * i + 2] = (uint8_t)((p[i] - lo) * sc);
    return write_bmp(path, g.data(), rows, cols);
}

}  // namespace bmpio
