// IoVti<Model>: dependency-free writer of the two VTK XML ImageData files the reference produces through VTK
// (src/lgca_io_vti.cpp:37-91 set-up, :113-144 write): `cell_res_<step>.vti` (points (dim_x+1) x (dim_y+1); cell data
// "Cell density", "Cell momentum" (2 components, AoS)) and `mean_res_<step>.vti` (points coarse_dim_x x coarse_dim_y;
// point data "Mean density", "Mean momentum").  Like the reference it wraps the lattice's host float arrays without
// copying (the pointers are captured once and must stay valid).  Encoding: raw appended binary, uncompressed
// (the reference asks VTK for LZ4; readers accept either -- compare array contents, not file bytes).
#ifndef LGCA_B200_HOST_IO_VTI_H_
#define LGCA_B200_HOST_IO_VTI_H_

#include <string>

#include "lattice.h"

namespace lgca {

template <Model model_>
class IoVti {
public:
    using LatticeType = Lattice<model_>;

    IoVti(LatticeType* lattice, const std::string scalars = "Cell density");

    void set_scalars(const std::string scalars) { m_scalars = scalars; }
    void update() {}  // nothing is cached: write() reads the live host arrays
    // writes <dir>cell_res_<step>.vti (unless the lattice has no per-cell fields) and <dir>mean_res_<step>.vti
    void write(const size_t step, const std::string dir = "./");

private:
    LatticeType* m_lattice;
    std::string  m_scalars;
    const Real*  m_cell_density;
    const Real*  m_cell_momentum;
    const Real*  m_mean_density;
    const Real*  m_mean_momentum;
};

} // namespace lgca

#endif
