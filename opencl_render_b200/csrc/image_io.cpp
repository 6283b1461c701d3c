// image_io.cpp -- image output of the render path: what the reference does with the three 16-bit
// planes after RaytraceAll returns (source/render.cpp:1372-1386, source/util/writebmp.cpp:124-177).
//
//   oclr_write_bmp    24-bit BMP, bottom-up BGR rows padded to 4 bytes, 54-byte header -- the layout writebmp3s emits.
//                     mode 0 (default): byte = value / 256, the conversion the plugin uses for the picture it shows
//                     (render.cpp:1381-1383); mode 1: byte = (unsigned char)value, the cast writebmp3s really performs
//                     (writebmp.cpp:136-141 -- it keeps the LOW byte, which turns smooth gradients into noise; kept only so a
//                     file can be compared byte for byte with one the reference wrote).
//   oclr_write_ppm16  binary PPM (P6, maxval 65535, big-endian samples): all 16 bits, no dependency.
//   oclr_write_png16  16-bit RGB PNG (zlib deflate, filter 0).
// Host code; no device involvement.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

#include "../../include/oclr_abi.h"

namespace {

void put32le(unsigned char* p, uint32_t v) {
    p[0] = (unsigned char)v;
    p[1] = (unsigned char)(v >> 8);
    p[2] = (unsigned char)(v >> 16);
    p[3] = (unsigned char)(v >> 24);
}
void put32be(unsigned char* p, uint32_t v) {
    p[0] = (unsigned char)(v >> 24);
    p[1] = (unsigned char)(v >> 16);
    p[2] = (unsigned char)(v >> 8);
    p[3] = (unsigned char)v;
}

bool png_chunk(FILE* f, const char type[4], const unsigned char* data, size_t len) {
    unsigned char head[8];
    put32be(head, (uint32_t)len);
    memcpy(head + 4, type, 4);
    uLong crc = crc32(0L, head + 4, 4);
    if (len) crc = crc32(crc, data, (uInt)len);
    unsigned char tail[4];
    put32be(tail, (uint32_t)crc);
    return fwrite(head, 1, 8, f) == 8 && (len == 0 || fwrite(data, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4;
}

}  // namespace

extern "C" {

int oclr_write_bmp(const char* path, cl_uint width, cl_uint height, const cl_ushort* red, const cl_ushort* green, const cl_ushort* blue,
                   int mode) {
    if (!path || !red || !green || !blue || width == 0 || height == 0 || (mode != 0 && mode != 1)) return 0;
    const size_t rowBytes = ((size_t)width * 3 + 3) & ~(size_t)3;
    const uint64_t fileSize = 54ull + (uint64_t)rowBytes * height;
    if (fileSize > 0xFFFFFFFFull) return 0;
    FILE* f = fopen(path, "wb");
    if (!f) return 0;
    unsigned char header[54] = {'B', 'M'};
    // mode 1 is writebmp3s byte for byte (writebmp.cpp:124-177): its bfSize ignores the row padding and biSizeImage stays 0
    put32le(header + 2, mode == 1 ? (uint32_t)(54ull + 3ull * width * height) : (uint32_t)fileSize);
    put32le(header + 10, 54);
    put32le(header + 14, 40);
    put32le(header + 18, width);
    put32le(header + 22, height);
    header[26] = 1;    // planes
    header[28] = 24;   // bits per pixel
    if (mode == 0) put32le(header + 34, (uint32_t)(rowBytes * height));
    bool ok = fwrite(header, 1, 54, f) == 54;
    std::vector<unsigned char> row(rowBytes, 0);
    for (cl_uint j = 0; ok && j < height; ++j) {
        const size_t src = (size_t)(height - 1 - j) * width;   // bottom-up
        for (cl_uint i = 0; i < width; ++i) {
            const cl_ushort r = red[src + i], g = green[src + i], b = blue[src + i];
            row[3 * (size_t)i + 2] = mode == 0 ? (unsigned char)(r / 256) : (unsigned char)r;
            row[3 * (size_t)i + 1] = mode == 0 ? (unsigned char)(g / 256) : (unsigned char)g;
            row[3 * (size_t)i + 0] = mode == 0 ? (unsigned char)(b / 256) : (unsigned char)b;
        }
        ok = fwrite(row.data(), 1, rowBytes, f) == rowBytes;
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? 1 : 0;
}

int oclr_write_ppm16(const char* path, cl_uint width, cl_uint height, const cl_ushort* red, const cl_ushort* green, const cl_ushort* blue) {
    if (!path || !red || !green || !blue || width == 0 || height == 0) return 0;
    FILE* f = fopen(path, "wb");
    if (!f) return 0;
    bool ok = fprintf(f, "P6\n%u %u\n65535\n", width, height) > 0;
    std::vector<unsigned char> row((size_t)width * 6);
    for (cl_uint j = 0; ok && j < height; ++j) {
        const size_t src = (size_t)j * width;
        for (cl_uint i = 0; i < width; ++i) {
            const cl_ushort v[3] = {red[src + i], green[src + i], blue[src + i]};
            for (int c = 0; c < 3; ++c) {
                row[6 * (size_t)i + 2 * c] = (unsigned char)(v[c] >> 8);
                row[6 * (size_t)i + 2 * c + 1] = (unsigned char)v[c];
            }
        }
        ok = fwrite(row.data(), 1, row.size(), f) == row.size();
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? 1 : 0;
}

int oclr_write_png16(const char* path, cl_uint width, cl_uint height, const cl_ushort* red, const cl_ushort* green, const cl_ushort* blue) {
    if (!path || !red || !green || !blue || width == 0 || height == 0) return 0;
    const size_t rowBytes = 1 + (size_t)width * 6;   // filter byte + RGB16 big-endian
    std::vector<unsigned char> raw(rowBytes * height);
    for (cl_uint j = 0; j < height; ++j) {
        unsigned char* row = raw.data() + rowBytes * j;
        row[0] = 0;
        const size_t src = (size_t)j * width;
        for (cl_uint i = 0; i < width; ++i) {
            const cl_ushort v[3] = {red[src + i], green[src + i], blue[src + i]};
            for (int c = 0; c < 3; ++c) {
                row[1 + 6 * (size_t)i + 2 * c] = (unsigned char)(v[c] >> 8);
                row[1 + 6 * (size_t)i + 2 * c + 1] = (unsigned char)v[c];
            }
        }
    }
    uLongf packedLen = compressBound((uLong)raw.size());
    std::vector<unsigned char> packed(packedLen);
    if (compress2(packed.data(), &packedLen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return 0;
    FILE* f = fopen(path, "wb");
    if (!f) return 0;
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    unsigned char ihdr[13];
    put32be(ihdr, width);
    put32be(ihdr + 4, height);
    ihdr[8] = 16;   // bit depth
    ihdr[9] = 2;    // colour type RGB
    ihdr[10] = ihdr[11] = ihdr[12] = 0;
    bool ok = fwrite(sig, 1, 8, f) == 8 && png_chunk(f, "IHDR", ihdr, 13) && png_chunk(f, "IDAT", packed.data(), packedLen) &&
              png_chunk(f, "IEND", nullptr, 0);
    ok = (fclose(f) == 0) && ok;
    return ok ? 1 : 0;
}

}  // extern "C"
