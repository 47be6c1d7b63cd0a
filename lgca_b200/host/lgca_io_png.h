// IoPng<Model>: colour-mapped PNG output of one lattice field -- the headless stand-in for the reference viewers'
// OUTPUT_FORMAT = "png" (apps/karman/karman_viewer.h:87; apps/pipe/pipe_viewer.cpp:171-181 writes res_<step>.png as
// a screenshot of the VTK render window).  What is kept of that picture is the field and its colours: the active
// scalars (default "Mean momentum", the viewers' default, apps/pipe/pipe_viewer.cpp:265; two-component arrays are
// mapped by magnitude like vtkLookupTable's default vector mode) over their own min..max range through VTK's rainbow
// table as the viewers configure it (apps/pipe/pipe_viewer.cpp:311-315: hue 2/3 -> 0, i.e. blue -> red, saturation
// and value 1, 256 entries), one pixel per (coarse) cell scaled up by an integer zoom, row 0 at the bottom.
// Window decorations (scalar bar, fonts, camera) are not reproduced: compare colours per cell, not file bytes.
// Dependency-free encoder: 8-bit RGB, stored (uncompressed) deflate blocks, CRC-32 / Adler-32 computed here.
#ifndef LGCA_B200_HOST_IO_PNG_H_
#define LGCA_B200_HOST_IO_PNG_H_

#include <cstdint>
#include <string>
#include <vector>

#include "lattice.h"

namespace lgca {

// 8-bit RGB image, rows top to bottom
bool write_png_rgb(const std::string& file, unsigned width, unsigned height, const uint8_t* rgb);

// vtkLookupTable semantics: 256 entries, hue 2/3 -> 0 over [lo, hi], values outside are clamped; lo == hi maps to entry 0
void colormap_blue_to_red(const float* values, size_t n, float lo, float hi, uint8_t* rgb);

template <Model model_>
class IoPng {
public:
    using LatticeType = Lattice<model_>;
    explicit IoPng(LatticeType* lattice, const std::string scalars = "Mean momentum", unsigned zoom = 1)
        : m_lattice(lattice), m_scalars(scalars), m_zoom(zoom ? zoom : 1) {}
    void set_scalars(const std::string scalars) { m_scalars = scalars; }
    // writes <dir>res_<step>.png; returns false when the active field is not available
    bool write(const size_t step, const std::string dir = "./");

private:
    LatticeType* m_lattice;
    std::string  m_scalars;
    unsigned     m_zoom;
};

} // namespace lgca

// C entry for tools and tests: maps `values` (components = 1, or 2 -> magnitude; row 0 = bottom) and writes the PNG.
// lo > hi = use the data range.  Returns 0 on success.
extern "C" int lgca_host_write_png(const char* file, unsigned width, unsigned height, const float* values, int components,
                                   float lo, float hi, unsigned zoom);

#endif
