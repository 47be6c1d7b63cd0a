// Lattice<Model>: the backend-independent part of an LGCA lattice -- sizing from (test case, Re, Ma, coarse
// graining radius), the host-side arrays in the reference's layouts, the cell-type (BC) painters, the particle
// initialisers and the forcing formulas.  Public/protected surface = the reference's abstract base class
// src/lattice.h:32-238, so an app written against it (apps/*/…_viewer.cpp, apps/periodic/main.cpp) compiles
// against this header unchanged and picks a backend by constructing OMP_Lattice or B200_Lattice.
//
// Host code only; written from scratch (behaviour per SURVEY.md 3.1-3.2, A.1, A.7, Appendix D).
#ifndef LGCA_B200_HOST_LATTICE_H_
#define LGCA_B200_HOST_LATTICE_H_

#include "lgca_bitset.h"
#include "lgca_common.h"
#include "lgca_models.h"

namespace lgca {

// float(rand()) / float(RAND_MAX): the reference's random_uniform(), src/utils.h:119-122
inline Real random_uniform() { return static_cast<Real>(std::rand()) / static_cast<Real>(RAND_MAX); }

template <Model model_>
class Lattice {
protected:
    using ModelDesc = ModelDescriptor<model_>;

    static constexpr unsigned int SPATIAL_DIM = 2;
    static constexpr unsigned int NUM_DIR     = ModelDesc::NUM_DIR;

    unsigned int m_dim_x = 0, m_dim_y = 0;
    size_t       m_num_cells = 0, m_num_nodes = 0, m_num_particles = 0;

    unsigned int m_coarse_graining_radius = 0, m_coarse_dim_x = 0, m_coarse_dim_y = 0;
    size_t       m_num_coarse_cells = 0;

    string m_test_case;

    const Real m_rho = 1.0; // density
    const Real m_c   = 1.0; // speed of sound

    Real m_Re = 0, m_Ma_s = 0; // Reynolds number, scaled Mach number
    Real m_d = 0;              // mean occupation number
    Real m_nu = 0;             // viscosity
    Real m_g = 0;              // Galilean breaking factor
    Real m_nu_s = 0;           // scaled viscosity
    Real m_c_s = 0;            // scaled sound speed
    Real m_u = 0;              // velocity
    char m_bf_dir = 0;         // body force direction: 'x', 'y' or 0

    int m_equilibrium_forcing = 0;

    // Host arrays (allocated and freed by the backend, like the reference: src/omp_lattice.cpp:457-489).
    CellType* m_cell_type_cpu = nullptr; // 0 fluid, 1 solid no-slip (bounce back), 2 solid slip (bounce forward)
    Bitset    m_node_state_cpu;          // bit (dir + cell*8) = occupation of direction dir in cell
    Bitset    m_node_state_out_cpu;      // snapshot for post-processing / visualisation
    Real*     m_cell_density_cpu  = nullptr;
    Real*     m_mean_density_cpu  = nullptr;
    Real*     m_cell_momentum_cpu = nullptr; // AoS [x0,y0,x1,y1,...]
    Real*     m_mean_momentum_cpu = nullptr;
    Bitset    m_rnd_cpu;                 // frozen per-cell chirality bits

    // explicit-dims extension (shapes the test-case formulas cannot produce, e.g. 65536 x 32768 "box")
    Lattice(const string test_case, unsigned int dim_x, unsigned int dim_y, const int coarse_graining_radius, char bf_dir);

public:
    Lattice(const string test_case, const Real Re, const Real Ma_s, const int coarse_graining_radius);
    virtual ~Lattice();

    void apply_cell_type_all(const CellType cell_type);
    void apply_boundary_cell_type_east(const CellType cell_type);
    void apply_boundary_cell_type_north(const CellType cell_type);
    void apply_boundary_cell_type_west(const CellType cell_type);
    void apply_boundary_cell_type_south(const CellType cell_type);

    void apply_bc_periodic();
    void apply_bc_reflecting(const string bounce_type);
    void apply_bc_pipe();
    void apply_bc_karman_vortex_street();

    void init_zero();
    void init_random();
    void init_single(const std::vector<size_t> occupied_nodes);
    void init_single_collision();
    void init_diffusion();

    // virtual here (a superset of the reference, where it is non-virtual): device backends count on the device
    virtual unsigned long get_n_particles();

    void print();
    void print_info();

    size_t get_equilibrium_forcing();
    size_t get_initial_forcing();

    virtual void setup_parallel() = 0;
    virtual void collide_and_propagate(const bool p = false) = 0;
    virtual std::vector<Real> get_mean_velocity() = 0;
    virtual void apply_body_force(const int forcing) = 0;
    virtual void post_process() = 0;

    virtual void copy_data_to_device();
    virtual void copy_data_from_device();
    virtual void copy_data_to_output_buffer();

    Real         nu_s()             const { return m_nu_s; }
    Real         c_s()              const { return m_c_s; }
    Real         u()                const { return m_u; }
    unsigned int dim_x()            const { return m_dim_x; }
    unsigned int dim_y()            const { return m_dim_y; }
    size_t       num_cells()        const { return m_num_cells; }
    unsigned int coarse_dim_x()     const { return m_coarse_dim_x; }
    unsigned int coarse_dim_y()     const { return m_coarse_dim_y; }
    size_t       num_coarse_cells() const { return m_num_coarse_cells; }

          Real* cell_density()        { assert(m_cell_density_cpu);  return m_cell_density_cpu; }
    const Real* cell_density()  const { assert(m_cell_density_cpu);  return m_cell_density_cpu; }
          Real* mean_density()        { assert(m_mean_density_cpu);  return m_mean_density_cpu; }
    const Real* mean_density()  const { assert(m_mean_density_cpu);  return m_mean_density_cpu; }
          Real* cell_momentum()       { assert(m_cell_momentum_cpu); return m_cell_momentum_cpu; }
    const Real* cell_momentum() const { assert(m_cell_momentum_cpu); return m_cell_momentum_cpu; }
          Real* mean_momentum()       { assert(m_mean_momentum_cpu); return m_mean_momentum_cpu; }
    const Real* mean_momentum() const { assert(m_mean_momentum_cpu); return m_mean_momentum_cpu; }

    Real cell_density(const int x, const int y) { assert(m_cell_density_cpu); return m_cell_density_cpu[y * m_dim_x + x]; }
    // (the reference indexes the coarse field with m_dim_x here, src/lattice.h:237 -- a bug nobody calls; fixed)
    Real mean_density(const int x, const int y) { assert(m_mean_density_cpu); return m_mean_density_cpu[y * m_coarse_dim_x + x]; }

    // raw host arrays for tools and tests (reference layouts)
    const uint8_t*  state_bytes() const { return m_node_state_cpu.ptr(); }
    uint8_t*        state_bytes()       { return m_node_state_cpu.ptr(); }
    const int32_t*  cell_types()  const { return reinterpret_cast<const int32_t*>(m_cell_type_cpu); }
    const uint8_t*  rnd_bits()    const { return m_rnd_cpu.ptr(); }
    char            bf_dir()      const { return m_bf_dir; }
    bool            has_cell_fields() const { return m_cell_density_cpu != nullptr && m_cell_momentum_cpu != nullptr; }
    // Backends that fill the per-cell host fields lazily (B200Options::lazy_cell_fields) bring them up to date with
    // the last post_process() here; writers call it before they read cell_density() / cell_momentum().
    virtual void    sync_cell_fields() {}
    const string&   test_case()   const { return m_test_case; }

private:
    void derive_physics();
    void finish_sizing(int coarse_graining_radius);
    void paint_edges(CellType t, bool south_north, bool east_west);
    bool draw_occupation() const;
};

} // namespace lgca

#endif
