#include "lgca_io_vti.h"

#include <cstdint>
#include <cstdio>
#include <sstream>

namespace lgca {

namespace {
struct Array {
    const char*  name;
    int          components;
    const float* data;
    uint64_t     count; // floats
};

// one ImageData piece with raw appended Float32 arrays (VTK XML format, header_type UInt64)
bool write_image(const std::string& file, unsigned ext_x, unsigned ext_y, bool cell_data, const std::string& active,
                 const Array* arrays, int n)
{
    FILE* f = std::fopen(file.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "<?xml version=\"1.0\"?>\n<VTKFile type=\"ImageData\" version=\"1.0\" byte_order=\"LittleEndian\" "
                    "header_type=\"UInt64\">\n");
    std::fprintf(f, "  <ImageData WholeExtent=\"0 %u 0 %u 0 0\" Origin=\"0 0 0\" Spacing=\"1 1 1\">\n", ext_x, ext_y);
    std::fprintf(f, "    <Piece Extent=\"0 %u 0 %u 0 0\">\n", ext_x, ext_y);
    const char* tag = cell_data ? "CellData" : "PointData";
    std::fprintf(f, "      <%s Scalars=\"%s\">\n", tag, active.c_str());
    uint64_t offset = 0;
    for (int i = 0; i < n; ++i) {
        std::fprintf(f, "        <DataArray type=\"Float32\" Name=\"%s\" NumberOfComponents=\"%d\" format=\"appended\" "
                        "offset=\"%llu\"/>\n", arrays[i].name, arrays[i].components, (unsigned long long)offset);
        offset += sizeof(uint64_t) + arrays[i].count * sizeof(float);
    }
    std::fprintf(f, "      </%s>\n    </Piece>\n  </ImageData>\n  <AppendedData encoding=\"raw\">\n   _", tag);
    for (int i = 0; i < n; ++i) {
        const uint64_t bytes = arrays[i].count * sizeof(float);
        std::fwrite(&bytes, sizeof(bytes), 1, f);
        std::fwrite(arrays[i].data, 1, bytes, f);
    }
    std::fprintf(f, "\n  </AppendedData>\n</VTKFile>\n");
    const bool ok = std::ferror(f) == 0;
    std::fclose(f);
    return ok;
}
} // namespace

template <Model model_>
IoVti<model_>::IoVti(LatticeType* lattice, const std::string scalars) : m_lattice(lattice), m_scalars(scalars)
{
    assert(m_lattice);
    // capture the host array pointers once, like the reference (they must stay stable)
    const LatticeType* cl = m_lattice;
    m_mean_density  = cl->mean_density();
    m_mean_momentum = cl->mean_momentum();
    m_cell_density  = nullptr;
    m_cell_momentum = nullptr;
}

template <Model model_>
void IoVti<model_>::write(const size_t step, const std::string dir)
{
    std::ostringstream cell_file, mean_file;
    cell_file << dir << "cell_res_" << step << ".vti";
    mean_file << dir << "mean_res_" << step << ".vti";
    const uint64_t n = m_lattice->num_cells(), nc = m_lattice->num_coarse_cells();

    if (m_lattice->has_cell_fields()) m_lattice->sync_cell_fields();
    const Real* rho = m_lattice->has_cell_fields() ? m_lattice->cell_density() : nullptr;
    const Real* mom = m_lattice->has_cell_fields() ? m_lattice->cell_momentum() : nullptr;
    if (rho && mom) {
        const Array a[2] = {{"Cell density", 1, rho, n}, {"Cell momentum", 2, mom, 2 * n}};
        const std::string active = m_scalars.rfind("Cell", 0) == 0 ? m_scalars : "Cell density";
        if (!write_image(cell_file.str(), m_lattice->dim_x(), m_lattice->dim_y(), true, active, a, 2))
            printf("ERROR in IoVti::write(): cannot write %s\n", cell_file.str().c_str());
    }
    if (nc > 0) {
        const Array a[2] = {{"Mean density", 1, m_mean_density, nc}, {"Mean momentum", 2, m_mean_momentum, 2 * nc}};
        const std::string active = m_scalars.rfind("Mean", 0) == 0 ? m_scalars : "Mean density";
        if (!write_image(mean_file.str(), m_lattice->coarse_dim_x() - 1, m_lattice->coarse_dim_y() - 1, false, active, a, 2))
            printf("ERROR in IoVti::write(): cannot write %s\n", mean_file.str().c_str());
    }
}

template class IoVti<Model::HPP>;
template class IoVti<Model::FHP_I>;
template class IoVti<Model::FHP_II>;
template class IoVti<Model::FHP_III>;

} // namespace lgca
