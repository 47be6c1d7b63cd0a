// src/b200_lattice.h for the REFERENCE tree -- the binding a keva92/lgca maintainer adds next to src/omp_lattice.h.
//
// Written against the reference's own, unmodified headers (src/lattice.h, src/lgca_bitset.h, src/lgca_models.h) and
// the C-ABI include/lgca_b200.h, nothing else: header-only, no CUDA at compile time, link with -llgca_b200.  An app
// selects the backend where it says `new OMP_Lattice<MODEL>(...)` (apps/pipe/pipe_viewer.cpp:46):
//
//     m_lattice = new B200_Lattice<MODEL>("pipe", m_Re, m_Ma, CG_RADIUS);            // one GPU
//     m_lattice = new B200_Lattice<MODEL>("pipe", m_Re, m_Ma, CG_RADIUS, /*n_gpus=*/8);  // row strips over 8 GPUs
//
// oracle/Makefile compiles exactly this file against /root/reference/src (target `ref`, binary
// oracle/_ref/ref_b200_app from integration/ref_app.cpp) and tests/test_integration_ref.py runs it: the reference's
// base class, BC painters, initialisers and forcing formulas + this backend reproduce the reference's hashes.
// (The repo's own host layer lgca_b200/host/ is the same class on a from-scratch base, with pinned mirrors and an
// explicit-dims constructor the reference base lacks.)
#ifndef LGCA_B200_LATTICE_REF_H_
#define LGCA_B200_LATTICE_REF_H_

#include <cstdio>
#include <cstdlib>
#include <deque>
#include <vector>

#include "lattice.h"

extern "C" {
#include "lgca_b200.h"
}

namespace lgca {

template<Model model_>
class B200_Lattice : public Lattice<model_> {

    lgca_b200_group* m_h = nullptr;        // the lattice on 1..n GPUs (lgca_b200_group_*)
    bool             m_on_device = false;  // host mirrors have been uploaded
    std::deque<int>  m_draws;              // rand() values drawn ahead for the body force, in stream order

    static int model_id() {
        return model_ == Model::HPP ? LGCA_B200_HPP : model_ == Model::FHP_I ? LGCA_B200_FHP_I
             : model_ == Model::FHP_II ? LGCA_B200_FHP_II : LGCA_B200_FHP_III;
    }
    void check(int rc, const char* where) {
        if (rc) { printf("ERROR in B200_Lattice::%s(): %s (code %d)\n", where, lgca_b200_last_error(), rc); fflush(stdout); abort(); }
    }
    void ensure_on_device() { if (!m_on_device) copy_data_to_device(); }
    bool coarse_ok() const {
        const unsigned int r = this->m_coarse_graining_radius;
        return r && this->m_dim_x % (2 * r) == 0 && this->m_dim_y % (2 * r) == 0 && this->m_dim_x >= 4 * r;
    }

    // host mirrors in the reference's layouts, as src/omp_lattice.cpp:457-471 allocates them
    void allocate_memory() {
        const size_t n = this->m_num_cells, nc = this->m_num_coarse_cells ? this->m_num_coarse_cells : 1;
        this->m_cell_type_cpu     = (CellType*)calloc(n, sizeof(CellType));
        this->m_cell_density_cpu  = (Real*)calloc(n, sizeof(Real));
        this->m_mean_density_cpu  = (Real*)calloc(nc, sizeof(Real));
        this->m_cell_momentum_cpu = (Real*)calloc(this->SPATIAL_DIM * n, sizeof(Real));
        this->m_mean_momentum_cpu = (Real*)calloc(this->SPATIAL_DIM * nc, sizeof(Real));
        this->m_node_state_cpu.resize(n * 8);
        this->m_node_state_out_cpu.resize(n * 8);   // kept for base-class readers; the snapshot itself lives on the device
        this->m_rnd_cpu.resize(n);
    }
    void free_memory() {
        free(this->m_cell_type_cpu); free(this->m_cell_density_cpu); free(this->m_mean_density_cpu);
        free(this->m_cell_momentum_cpu); free(this->m_mean_momentum_cpu);
        this->m_cell_type_cpu = NULL;
        this->m_cell_density_cpu = this->m_mean_density_cpu = this->m_cell_momentum_cpu = this->m_mean_momentum_cpu = NULL;
    }

public:

    B200_Lattice(const string test_case, const Real Re, const Real Ma_s, const int coarse_graining_radius,
                 const int n_gpus = 1, const int k_fuse = 0)
        : Lattice<model_>(test_case, Re, Ma_s, coarse_graining_radius)
    {
        allocate_memory();
        this->m_rnd_cpu.fill_random();   // same place in the rand() stream as src/omp_lattice.cpp:84
        lgca_b200_config c = lgca_b200_config();
        c.model = model_id(); c.dim_x = this->m_dim_x; c.dim_y = this->m_dim_y;
        c.cg_radius = coarse_ok() ? this->m_coarse_graining_radius : 0;   // the 21 x 10 "collision" demo has no valid coarse grid
        c.bf_dir = this->m_bf_dir; c.k_fuse = k_fuse;
        check(lgca_b200_group_create(&c, n_gpus, NULL, &m_h), "B200_Lattice");
    }
    virtual ~B200_Lattice() { lgca_b200_group_destroy(m_h); free_memory(); }

    void setup_parallel() {
        lgca_b200_info i;
        check(lgca_b200_group_get_info(m_h, &i), "setup_parallel");
        printf("B200 configuration parameters: %u bit-planes of %u x %u words, %d fused steps per pass.\n\n",
               i.num_planes, i.y_rows, i.words_per_row, i.k_fuse);
    }
    void copy_data_to_device() {
        check(lgca_b200_group_upload(m_h, this->m_node_state_cpu.ptr(), (const int32_t*)this->m_cell_type_cpu, this->m_rnd_cpu.ptr()),
              "copy_data_to_device");
        m_on_device = true;
    }
    void copy_data_from_device() { ensure_on_device(); check(lgca_b200_group_download(m_h, this->m_node_state_cpu.ptr()), "copy_data_from_device"); }
    void collide_and_propagate(const bool /*p: unused by the reference's live backend too*/) {
        ensure_on_device();
        check(lgca_b200_group_step(m_h, 1), "collide_and_propagate");
    }
    void collide_and_propagate_n(const int n) { ensure_on_device(); check(lgca_b200_group_step(m_h, n), "collide_and_propagate"); }
    void copy_data_to_output_buffer() { ensure_on_device(); check(lgca_b200_group_snapshot(m_h), "copy_data_to_output_buffer"); }
    void post_process() {
        ensure_on_device();
        const bool c = coarse_ok();
        check(lgca_b200_group_post_process(m_h, this->m_cell_density_cpu, this->m_cell_momentum_cpu, c ? this->m_mean_density_cpu : NULL,
                                           c ? this->m_mean_momentum_cpu : NULL, /*exact_order=*/1), "post_process");
    }
    // sequential float32 sums over the per-cell host fields of the last post_process(), one thread: the only order that
    // reproduces the digits of src/omp_lattice.cpp:508-557
    std::vector<Real> get_mean_velocity() {
        // order-exact on the device (the one-thread float32 sums of src/omp_lattice.cpp:508-557, see csrc/lgca_mv.cu)
        std::vector<Real> v(this->SPATIAL_DIM, 0.0);
        ensure_on_device();
        float out[2];
        check(lgca_b200_group_mean_velocity_exact(m_h, out), "get_mean_velocity");
        v[0] = out[0]; v[1] = out[1];
        return v;
    }
    // rand() values are drawn ahead into a FIFO, consumed in order by the exact device body force, leftovers are kept:
    // the stream position after the call equals the reference's (src/omp_lattice.cpp:254-346)
    void apply_body_force(const int forcing) {
        ensure_on_device();
        const size_t it_max = 2 * this->m_num_cells;
        size_t it = 0;
        long remaining = (long)(unsigned int)forcing;
        bool first = true;
        std::vector<int32_t> batch;
        while ((first || remaining > 0) && it < it_max) {
            size_t want = (size_t)(remaining > 0 ? remaining : 1) * 12 + 256;
            if (want > it_max - it) want = it_max - it;
            while (m_draws.size() < want) m_draws.push_back(rand());
            batch.assign(m_draws.begin(), m_draws.begin() + want);
            size_t consumed = 0; uint32_t reverted = 0;
            check(lgca_b200_group_body_force(m_h, (int)remaining, batch.data(), want, &consumed, &reverted), "apply_body_force");
            m_draws.erase(m_draws.begin(), m_draws.begin() + consumed);
            it += consumed; remaining -= reverted; first = false;
            if (!consumed) break;
        }
    }
    void synchronize() { check(lgca_b200_group_sync(m_h), "synchronize"); }
    // raw host state for tools (the reference keeps it protected)
    const unsigned char* state_bytes() const { return this->m_node_state_cpu.ptr(); }
};

} // namespace lgca

#endif
