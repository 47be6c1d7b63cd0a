// Host-side logic of Lattice<Model> (see lattice.h).  Behaviour follows the reference's src/lattice.cpp
// (sizing :29-157, painters :221-395, initialisers :198-217/:283-308/:465-496, forcing :445-461); the float/double
// mix of every formula is kept because lattice sizes and the rand() stream positions depend on it.
#include "lattice.h"

#include <cstring>

namespace lgca {

namespace {
enum class Case { PIPE, KARMAN, COLLISION, DIFFUSION, PERIODIC, BOX, INVALID };
Case parse_case(const string& s)
{
    if (s == "pipe") return Case::PIPE;
    if (s == "karman") return Case::KARMAN;
    if (s == "collision") return Case::COLLISION;
    if (s == "diffusion") return Case::DIFFUSION;
    if (s == "periodic") return Case::PERIODIC;
    if (s == "box") return Case::BOX;
    return Case::INVALID;
}
} // namespace

template <Model model_>
void Lattice<model_>::derive_physics()
{
    // all members are float; literals are double, exactly as in the reference (src/lattice.cpp:52-67)
    m_d    = m_rho / NUM_DIR;
    m_nu   = 1.0 / 12.0 * 1.0 / (m_d * pow((1.0 - m_d), 3.0)) - 1.0 / 8.0;
    m_g    = SPATIAL_DIM / (SPATIAL_DIM + 2.0) * (1.0 - 2.0 * m_d) / (1.0 - m_d);
    m_nu_s = m_nu / m_g;
    m_c_s  = m_c / sqrt((double)SPATIAL_DIM);
    m_u    = m_Ma_s * m_c_s;
}

template <Model model_>
void Lattice<model_>::finish_sizing(int cg)
{
    assert(m_dim_x > 0 && m_dim_y > 0);
    if (model_ != Model::HPP) assert(m_dim_y % 2 == 0);
    m_num_cells = (size_t)m_dim_x * m_dim_y; // 64-bit here (the reference overflows at 2^32 cells, src/lattice.cpp:144)
    m_num_nodes = m_num_cells * NUM_DIR;
    m_num_particles = 0;
    assert(cg > 0);
    m_coarse_graining_radius = (unsigned)cg;
    m_coarse_dim_x = m_dim_x / (2 * m_coarse_graining_radius);
    m_coarse_dim_y = m_dim_y / (2 * m_coarse_graining_radius);
    assert(m_dim_x % (2 * m_coarse_graining_radius) == 0);
    assert(m_dim_y % (2 * m_coarse_graining_radius) == 0);
    m_num_coarse_cells = (size_t)m_coarse_dim_x * m_coarse_dim_y;
}

template <Model model_>
Lattice<model_>::Lattice(const string test_case, const Real Re, const Real Ma_s, const int cg)
{
    m_test_case = test_case;
    assert(Re > 1.0e-06);
    assert(Ma_s > 1.0e-06);
    m_Re   = Re;
    m_Ma_s = Ma_s;
    derive_physics();

    const Case c = parse_case(test_case);
    switch (c) {
    case Case::PIPE:      m_dim_y = (int)((Re * m_nu_s) / m_u + 0.5); break;
    case Case::KARMAN:    { Real diameter = (Re * m_nu_s) / m_u; m_dim_y = (int)(3.0 * diameter + 0.5); break; }
    case Case::COLLISION: m_dim_y = 8; break;
    case Case::DIFFUSION:
    case Case::PERIODIC:
    case Case::BOX:       m_dim_y = (int)Re; break;
    default:
        printf("ERROR in Lattice::Lattice(): Invalid test case %s.\n", test_case.c_str());
        abort();
    }
    // round up to the next multiple of 2*cg; always adds between 1 and 2*cg rows (src/lattice.cpp:100)
    m_dim_y += (2 * cg) - (m_dim_y % (2 * cg));

    const bool wide = (c == Case::PIPE || c == Case::KARMAN || c == Case::COLLISION);
    m_dim_x = wide ? 2 * m_dim_y : m_dim_y;
    if (c == Case::COLLISION) m_dim_x++;
    m_bf_dir = (c == Case::PIPE || c == Case::KARMAN) ? 'x' : 0;

    finish_sizing(cg);
    print_info();
}

template <Model model_>
Lattice<model_>::Lattice(const string test_case, unsigned int dim_x, unsigned int dim_y, const int cg, char bf_dir)
{
    m_test_case = test_case;
    m_Re   = 80.0;
    m_Ma_s = 0.2;
    derive_physics();
    m_dim_x  = dim_x;
    m_dim_y  = dim_y;
    m_bf_dir = bf_dir;
    finish_sizing(cg);
    print_info();
}

template <Model model_>
Lattice<model_>::~Lattice() {}

template <Model model_>
void Lattice<model_>::init_zero()
{
    m_node_state_cpu.reset();
}

template <Model model_>
void Lattice<model_>::print()
{
    m_node_state_cpu.print();
}

template <Model model_>
unsigned long Lattice<model_>::get_n_particles()
{
    // counts all 8 bits of every cell byte, like the reference (src/lattice.cpp:180-195)
    m_num_particles = m_node_state_cpu.count();
    return m_num_particles;
}

template <Model model_>
bool Lattice<model_>::draw_occupation() const
{
    return random_uniform() > (1.0 - (1.0 / NUM_DIR));
}

template <Model model_>
void Lattice<model_>::init_random()
{
    // serial, cells ascending, directions ascending, FLUID cells only: this fixes the rand() stream order
    // (the reference's `omp parallel for` here is only deterministic at one thread)
    uint8_t* s = m_node_state_cpu.ptr();
    for (size_t cell = 0; cell < m_num_cells; ++cell) {
        if (m_cell_type_cpu[cell] != CellType::FLUID) continue;
        uint8_t b = s[cell];
        for (unsigned dir = 0; dir < NUM_DIR; ++dir) {
            if (draw_occupation()) b |= (uint8_t)(1u << dir);
            else b &= (uint8_t)~(1u << dir);
        }
        s[cell] = b;
    }
}

template <Model model_>
void Lattice<model_>::init_diffusion()
{
    const int  center_x = m_dim_x / 2;
    const int  center_y = m_dim_y / 2;
    const Real diameter = m_dim_y / 4;
    uint8_t* s = m_node_state_cpu.ptr();
    for (size_t cell = 0; cell < m_num_cells; ++cell) {
        const int pos_x = cell % m_dim_x;
        const int pos_y = cell / m_dim_x;
        const Real dist = sqrt(pow((pos_x - center_x), 2.0) + pow((pos_y - center_y), 2.0));
        if (m_cell_type_cpu[cell] == CellType::FLUID && dist < (diameter / 2.0)) {
            uint8_t b = s[cell];
            for (unsigned dir = 0; dir < NUM_DIR; ++dir) {
                if (draw_occupation()) b |= (uint8_t)(1u << dir);
                else b &= (uint8_t)~(1u << dir);
            }
            s[cell] = b;
        }
    }
}

template <Model model_>
void Lattice<model_>::init_single(const std::vector<size_t> occupied_nodes)
{
    init_zero();
    for (size_t n = 0; n < occupied_nodes.size(); ++n) m_node_state_cpu[occupied_nodes[n]] = true;
}

template <Model model_>
void Lattice<model_>::init_single_collision()
{
    // two particles on a collision course in the middle row (node index = dir + cell*8)
    const int inverse_dir = ModelDesc::INV_DIR[0];
    std::vector<size_t> nodes;
    nodes.push_back((size_t)(m_dim_x * m_dim_y / 2 + 1) * 8);
    nodes.push_back((size_t)(m_dim_x * m_dim_y / 2 + 5) * 8 + inverse_dir);
    init_single(nodes);
}

// ---- cell-type painters ------------------------------------------------------------------------------------
template <Model model_>
void Lattice<model_>::apply_cell_type_all(const CellType t)
{
    for (size_t cell = 0; cell < m_num_cells; ++cell) m_cell_type_cpu[cell] = t;
}
template <Model model_>
void Lattice<model_>::apply_boundary_cell_type_east(const CellType t)
{
    for (size_t cell = m_dim_x - 1; cell < m_num_cells; cell += m_dim_x) m_cell_type_cpu[cell] = t;
}
template <Model model_>
void Lattice<model_>::apply_boundary_cell_type_north(const CellType t)
{
    for (size_t cell = m_num_cells - m_dim_x; cell < m_num_cells; ++cell) m_cell_type_cpu[cell] = t;
}
template <Model model_>
void Lattice<model_>::apply_boundary_cell_type_west(const CellType t)
{
    for (size_t cell = 0; cell < m_num_cells; cell += m_dim_x) m_cell_type_cpu[cell] = t;
}
template <Model model_>
void Lattice<model_>::apply_boundary_cell_type_south(const CellType t)
{
    for (size_t cell = 0; cell < m_dim_x; ++cell) m_cell_type_cpu[cell] = t;
}

template <Model model_>
void Lattice<model_>::paint_edges(CellType t, bool south_north, bool east_west)
{
    apply_cell_type_all(CellType::FLUID);
    // same painting order as the reference (east, north, west, south): corners end up with the last writer,
    // which is immaterial because one call paints a single type
    if (east_west) apply_boundary_cell_type_east(t);
    if (south_north) apply_boundary_cell_type_north(t);
    if (east_west) apply_boundary_cell_type_west(t);
    if (south_north) apply_boundary_cell_type_south(t);
}

template <Model model_>
void Lattice<model_>::apply_bc_periodic()
{
    paint_edges(CellType::FLUID, false, false);
}

template <Model model_>
void Lattice<model_>::apply_bc_pipe()
{
    paint_edges(CellType::SOLID_NO_SLIP, true, false);
}

template <Model model_>
void Lattice<model_>::apply_bc_reflecting(const string bounce_type)
{
    CellType t;
    if (bounce_type == "back") t = CellType::SOLID_NO_SLIP;
    else if (bounce_type == "forward") t = CellType::SOLID_SLIP;
    else {
        printf("ERROR in apply_bc_reflecting(): Invalid bounce type %s.\n", bounce_type.c_str());
        abort();
    }
    paint_edges(t, true, true);
}

template <Model model_>
void Lattice<model_>::apply_bc_karman_vortex_street()
{
    apply_bc_pipe();
    // cylinder: centre (dim_x/6, dim_y/2) -- the reference's `1 / 10 * m_dim_y` offset is integer zero --,
    // diameter float(dim_y / 3) with integer division, distance rounded to float, strict '<'
    const int  center_x = m_dim_x / 6;
    const int  center_y = m_dim_y / 2 + 1 / 10 * m_dim_y;
    const Real diameter = m_dim_y / 3;
    for (size_t cell = 0; cell < m_num_cells; ++cell) {
        const int pos_x = cell % m_dim_x;
        const int pos_y = cell / m_dim_x;
        const Real dist = sqrt(pow((pos_x - center_x), 2.0) + pow((pos_y - center_y), 2.0));
        if (dist < (diameter / 2.0)) m_cell_type_cpu[cell] = CellType::SOLID_NO_SLIP;
    }
}

template <Model model_>
void Lattice<model_>::print_info()
{
    printf("Parameter for test case \"%s\":\n\n", m_test_case.c_str());
    printf("Reynolds number             Re   = %10.2f\n", m_Re);
    printf("Mach number                 Ma   = %10.2f\n", m_Ma_s);
    printf("Density                     rho  = %10.2f\n", m_rho);
    printf("Velocity                    u    = %10.2f\n", m_u);
    printf("Sound speed                 c    = %10.2f\n", m_c);
    printf("Scaled sound speed          c_s  = %10.2f\n", m_c_s);
    printf("Viscosity                   nu   = %10.2f\n", m_nu);
    printf("Scaled viscosity            nu_s = %10.2f\n", m_nu_s);
    printf("Galilean breaking factor    g    = %10.2f\n\n", m_g);
    printf("Number of cells in x direction: %d\n", m_dim_x);
    printf("Number of cells in y direction: %d\n\n", m_dim_y);
    printf("Number of coarse cells in x direction: %d\n", m_coarse_dim_x);
    printf("Number of coarse cells in y direction: %d\n\n", m_coarse_dim_y);
}

// hooks for device backends (empty in the base class, like the reference: src/lattice.cpp:422-435)
template <Model model_> void Lattice<model_>::copy_data_to_device() {}
template <Model model_> void Lattice<model_>::copy_data_from_device() {}

template <Model model_>
void Lattice<model_>::copy_data_to_output_buffer()
{
    m_node_state_out_cpu.copy(m_node_state_cpu);
}

template <Model model_>
size_t Lattice<model_>::get_initial_forcing()
{
    return (size_t)(0.01 * m_num_cells);
}

template <Model model_>
size_t Lattice<model_>::get_equilibrium_forcing()
{
    const Real forcing = (8.0 * m_nu_s * m_Ma_s * m_c_s) / pow((Real)m_dim_y, 2.0);
    return ceil(0.5 * m_num_cells * forcing);
}

template class Lattice<Model::HPP>;
template class Lattice<Model::FHP_I>;
template class Lattice<Model::FHP_II>;
template class Lattice<Model::FHP_III>;

} // namespace lgca
