#include "lgca_io_png.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <sstream>

namespace lgca {

namespace {

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n)
{
    static uint32_t table[256];
    static bool     ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return crc;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x)
{
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

bool write_chunk(FILE* f, const char type[4], const std::vector<uint8_t>& data)
{
    std::vector<uint8_t> head;
    put_be32(head, (uint32_t)data.size());
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32_update(0xFFFFFFFFu, head.data() + 4, 4);
    if (!data.empty()) crc = crc32_update(crc, data.data(), data.size());
    std::vector<uint8_t> tail;
    put_be32(tail, crc ^ 0xFFFFFFFFu);
    return std::fwrite(head.data(), 1, head.size(), f) == head.size() &&
           (data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size()) &&
           std::fwrite(tail.data(), 1, tail.size(), f) == tail.size();
}

// hue in [0, 1), full saturation and value -> 8-bit RGB (the six-sector HSV cone, as vtkMath::HSVToRGB)
void hue_to_rgb(double h, uint8_t* rgb)
{
    const double onethird = 1.0 / 3.0, onesixth = 1.0 / 6.0, twothird = 2.0 / 3.0, fivesixth = 5.0 / 6.0;
    double r, g, b;
    if (h > onesixth && h <= onethird)      { g = 1.0; r = (onethird - h) / onesixth; b = 0.0; }
    else if (h > onethird && h <= 0.5)      { g = 1.0; b = (h - onethird) / onesixth; r = 0.0; }
    else if (h > 0.5 && h <= twothird)      { b = 1.0; g = (twothird - h) / onesixth; r = 0.0; }
    else if (h > twothird && h <= fivesixth){ b = 1.0; r = (h - twothird) / onesixth; g = 0.0; }
    else if (h > fivesixth && h <= 1.0)     { r = 1.0; b = (1.0 - h) / onesixth; g = 0.0; }
    else                                    { r = 1.0; g = h / onesixth; b = 0.0; }
    rgb[0] = (uint8_t)std::lround(r * 255.0);
    rgb[1] = (uint8_t)std::lround(g * 255.0);
    rgb[2] = (uint8_t)std::lround(b * 255.0);
}

} // namespace

bool write_png_rgb(const std::string& file, unsigned width, unsigned height, const uint8_t* rgb)
{
    FILE* f = std::fopen(file.c_str(), "wb");
    if (!f) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    bool ok = std::fwrite(sig, 1, 8, f) == 8;
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width);
    put_be32(ihdr, height);
    const uint8_t rest[5] = {8, 2, 0, 0, 0}; // bit depth 8, colour type 2 (RGB), deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), rest, rest + 5);
    ok = ok && write_chunk(f, "IHDR", ihdr);
    // raw scanlines: filter byte 0 + RGB triples
    const size_t stride = (size_t)width * 3 + 1;
    std::vector<uint8_t> raw(stride * height);
    for (unsigned y = 0; y < height; ++y) {
        raw[y * stride] = 0;
        std::copy(rgb + (size_t)y * width * 3, rgb + (size_t)(y + 1) * width * 3, raw.begin() + y * stride + 1);
    }
    // zlib stream of stored blocks (<= 65535 bytes each) + Adler-32
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    size_t pos = 0;
    do {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(n & 0xFF)); z.push_back((uint8_t)(n >> 8));
        z.push_back((uint8_t)(~n & 0xFF)); z.push_back((uint8_t)((~n >> 8) & 0xFF));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n; ++i) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        pos += n;
    } while (pos < raw.size());
    put_be32(z, (b << 16) | a);
    ok = ok && write_chunk(f, "IDAT", z) && write_chunk(f, "IEND", std::vector<uint8_t>());
    ok = ok && std::ferror(f) == 0;
    std::fclose(f);
    return ok;
}

void colormap_blue_to_red(const float* values, size_t n, float lo, float hi, uint8_t* rgb)
{
    uint8_t table[256][3];
    for (int i = 0; i < 256; ++i) hue_to_rgb((2.0 / 3.0) * (1.0 - i / 255.0), table[i]); // hue 2/3 (blue) -> 0 (red)
    const double scale = hi > lo ? 256.0 / ((double)hi - (double)lo) : 0.0;
    for (size_t k = 0; k < n; ++k) {
        double x = ((double)values[k] - (double)lo) * scale;
        int    i = x != x ? 0 : (int)std::floor(x);
        i = std::max(0, std::min(255, i));
        rgb[3 * k] = table[i][0]; rgb[3 * k + 1] = table[i][1]; rgb[3 * k + 2] = table[i][2];
    }
}

namespace {
bool map_and_write(const std::string& file, unsigned w, unsigned h, const float* v, int comps, float lo, float hi, unsigned zoom)
{
    const size_t n = (size_t)w * h;
    std::vector<float> s(n);
    for (size_t i = 0; i < n; ++i)
        s[i] = comps == 2 ? std::sqrt(v[2 * i] * v[2 * i] + v[2 * i + 1] * v[2 * i + 1]) : v[i];
    if (lo > hi && n) { lo = *std::min_element(s.begin(), s.end()); hi = *std::max_element(s.begin(), s.end()); }
    std::vector<uint8_t> cell(3 * n);
    colormap_blue_to_red(s.data(), n, lo, hi, cell.data());
    zoom = zoom ? zoom : 1;
    const unsigned W = w * zoom, H = h * zoom;
    std::vector<uint8_t> img((size_t)3 * W * H);
    for (unsigned y = 0; y < H; ++y) {
        const unsigned cy = h - 1 - y / zoom; // row 0 of the lattice is the bottom of the picture
        for (unsigned x = 0; x < W; ++x) {
            const uint8_t* c = &cell[3 * ((size_t)cy * w + x / zoom)];
            uint8_t* o = &img[3 * ((size_t)y * W + x)];
            o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
        }
    }
    return write_png_rgb(file, W, H, img.data());
}
} // namespace

template <Model model_>
bool IoPng<model_>::write(const size_t step, const std::string dir)
{
    std::ostringstream file;
    file << dir << "res_" << step << ".png";
    const LatticeType* cl = m_lattice;
    const bool mean = m_scalars.rfind("Mean", 0) == 0;
    const bool vec  = m_scalars.find("momentum") != std::string::npos;
    if (!mean && !cl->has_cell_fields()) return false;
    if (!mean) m_lattice->sync_cell_fields();
    if (mean && cl->num_coarse_cells() == 0) return false;
    const unsigned w = mean ? cl->coarse_dim_x() : cl->dim_x(), h = mean ? cl->coarse_dim_y() : cl->dim_y();
    const Real* v = mean ? (vec ? cl->mean_momentum() : cl->mean_density()) : (vec ? cl->cell_momentum() : cl->cell_density());
    if (!map_and_write(file.str(), w, h, v, vec ? 2 : 1, 1.0f, 0.0f, m_zoom)) {
        printf("ERROR in IoPng::write(): cannot write %s\n", file.str().c_str());
        return false;
    }
    return true;
}

template class IoPng<Model::HPP>;
template class IoPng<Model::FHP_I>;
template class IoPng<Model::FHP_II>;
template class IoPng<Model::FHP_III>;

} // namespace lgca

extern "C" int lgca_host_write_png(const char* file, unsigned width, unsigned height, const float* values, int components,
                                   float lo, float hi, unsigned zoom)
{
    if (!file || !values || width == 0 || height == 0 || (components != 1 && components != 2)) return -1;
    return lgca::map_and_write(file, width, height, values, components, lo, hi, zoom) ? 0 : -2;
}
