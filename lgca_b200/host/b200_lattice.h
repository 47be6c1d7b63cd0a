// B200_Lattice<Model>: the B200 backend behind the reference's Lattice<Model> interface -- the class an app
// constructs where it used to say `new OMP_Lattice<MODEL>(...)` (reference: src/omp_lattice.h:29-78,
// apps/pipe/pipe_viewer.cpp:46).  All lattice arithmetic happens in liblgca_b200.so through the C-ABI
// (include/lgca_b200.h); this class only owns the host mirrors and keeps the reference's call semantics:
//
//   * ctor allocates, then draws the chirality bits (rand() stream order of src/omp_lattice.cpp:74-88);
//   * BC painters / initialisers of the base class work on the host mirrors; the first device-facing call
//     (collide_and_propagate, apply_body_force, copy_data_to_output_buffer, post_process, get_n_particles)
//     uploads them automatically, so apps that never call copy_data_to_device() still work;
//   * post_process() fills the same four host float arrays (stable pointers, like IoVti expects);
//   * get_mean_velocity() is the reference's sequential float32 loop over those host fields (bit-exact);
//   * apply_body_force() feeds the caller's rand() stream, in order, to the exact device body force;
//   * errors print "ERROR in ..." and abort(), like the reference;
//   * B200Options::n_gpus > 1 spreads the lattice over several GPUs of the box (row strips + halo ring behind
//     lgca_b200_group_*): same calls, same host arrays, identical results.
#ifndef LGCA_B200_HOST_B200_LATTICE_H_
#define LGCA_B200_HOST_B200_LATTICE_H_

#include <thread>

#include "lattice.h"

#include <vector>

struct lgca_b200_group;

namespace lgca {

struct B200Options {
    int  device  = 0;     // CUDA device ordinal (single GPU)
    int  n_gpus  = 1;     // > 1: the lattice is cut into row strips over n_gpus devices of this box (halo ring over NVLink)
    std::vector<int> devices;  // explicit device ordinals of the strips (default: device, device+1, ...)
    int  k_fuse  = 0;     // time steps fused per HBM pass (0 = library default)
    bool cell_fields = true;  // produce per-cell density/momentum in post_process (needs 12 B/cell host+device)
    bool lazy_cell_fields = false; // post_process() leaves the per-cell host fields alone (12 B/cell over PCIe per call);
                                   // sync_cell_fields() computes them from the same output buffer when a writer asks
    bool exact_post  = true;  // reference summation order for the coarse momentum-y means
    bool bulk_rand = false;      // draw the body-force rand() values in bulk: the state of glibc's default generator (TYPE_3
                                 // additive feedback, r[i] = r[i-31] + r[i-3]) is borrowed through initstate()/setstate(),
                                 // advanced by the same recurrence without the per-call lock, and handed back -- the stream and
                                 // every later rand() call are unchanged (self-checked against rand() on first use; falls back
                                 // to rand() for any other generator type).  Same caveat as prefetch_draws.
    bool prefetch_draws = false; // get_mean_velocity() starts a helper thread that draws the rand() values the following
                                 // apply_body_force() is expected to consume (glibc rand(): 6-20 ns per draw, ~8 draws per
                                 // reverted particle) while the mean velocity is computed.  The draws enter the same FIFO in
                                 // the same order; only for callers whose other threads do not call rand() meanwhile.
};

template <Model model_>
class B200_Lattice : public Lattice<model_> {
public:
    B200_Lattice(const string test_case, const Real Re, const Real Ma_s, const int coarse_graining_radius,
                 const B200Options& opt = B200Options());
    // explicit-dims extension (e.g. the 65536 x 32768 box of BASELINE config C4)
    B200_Lattice(const string test_case, unsigned int dim_x, unsigned int dim_y, const int coarse_graining_radius,
                 char bf_dir, const B200Options& opt = B200Options());
    virtual ~B200_Lattice();

    void setup_parallel() override;
    void collide_and_propagate(const bool p = false) override;
    void collide_and_propagate_n(int n_steps);           // n fused updates in one call (extension)
    std::vector<Real> get_mean_velocity() override;
    void apply_body_force(const int forcing) override;
    void post_process() override;
    void sync_cell_fields() override;
    void copy_data_to_device() override;
    void copy_data_from_device() override;
    void copy_data_to_output_buffer() override;
    unsigned long get_n_particles() override;

    void   synchronize();
    double timed_steps(int n_steps);                      // device time [ms] of n updates (CUDA events)
    lgca_b200_group* handle() { return m_h; }
    int  n_gpus() const { return m_opt.n_gpus; }

private:
    void allocate_memory();
    void free_memory();
    void create_device_lattice();
    void ensure_on_device();
    void fail(const char* where, int rc);

    B200Options        m_opt;
    lgca_b200_group*   m_h = nullptr;     // one lattice on n_gpus devices (n_gpus == 1: a plain whole-lattice handle)
    bool               m_on_device = false;   // host mirrors have been uploaded
    bool               m_cell_fields_stale = false; // lazy mode: the last post_process() skipped the per-cell fields
    std::vector<int32_t> m_draws;             // rand() values drawn ahead for the body force, in stream order;
    size_t             m_draw_head = 0;       // the unconsumed ones are m_draws[m_draw_head ..]
    size_t             draws_pending() const { return m_draws.size() - m_draw_head; }
    std::thread        m_prefetch;            // helper drawing ahead into m_draws (prefetch_draws)
    size_t             m_last_consumed = 0;   // draws the last apply_body_force() consumed
    void               join_prefetch() { if (m_prefetch.joinable()) m_prefetch.join(); }
    void               draw_until(size_t pending);  // top the FIFO up to `pending` unconsumed draws (rand() or bulk)
    double             m_draws_per_hit = 8.0; // running estimate used to size the draw-ahead
};

} // namespace lgca

#endif
