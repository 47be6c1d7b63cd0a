// B200_Lattice<Model> -- see b200_lattice.h.  Host C++ on top of the C-ABI; no lattice arithmetic here.
#include "b200_lattice.h"
#include <chrono>
#include <cstdio>

#include <algorithm>
#include <cstring>

#include "../../include/lgca_b200.h"

namespace lgca {

// ---- bulk rand() ----------------------------------------------------------------------------------------------------
// glibc's rand() is random(): a lock, then one step of the TYPE_3 additive feedback generator on its default state table
// (stdlib/random_r.c: `val = *fptr += *rptr; result = val >> 1`, degree 31, separation 3).  initstate() / setstate() are the
// documented way to move that generator between state arrays: initstate(seed, mine, n) makes it use `mine` and returns
// the OLD array, whose first word then holds 5 * rear_index + type.  While the generator is parked on the dummy array the
// default table is advanced here by the same recurrence, the rear index is written back and setstate() returns the
// generator to it: the stream continues exactly where bulk generation stopped.
bool bulk_rand_fill(std::vector<int32_t>& out, size_t n)
{
    static char dummy[128];
    char* old = initstate(1u, dummy, sizeof(dummy));
    if (!old) return false;
    int32_t* st = reinterpret_cast<int32_t*>(old);
    const int type = st[0] % 5, rear0 = st[0] / 5;
    bool ok = type == 3 && rear0 >= 0 && rear0 < 31;
    if (ok) {
        uint32_t* state = reinterpret_cast<uint32_t*>(st + 1);
        int r = rear0, f = (rear0 + 3) % 31;
        const size_t base = out.size();
        out.resize(base + n);
        int32_t* dst = out.data() + base;
        for (size_t i = 0; i < n; ++i) {
            const uint32_t val = state[f] += state[r];
            dst[i] = (int32_t)(val >> 1);
            if (++f == 31) f = 0;
            if (++r == 31) r = 0;
        }
        st[0] = 5 * r + type;
    }
    setstate(old);
    return ok;
}

// first use: the bulk generator must reproduce rand() itself -- 32 values computed on a COPY of the state are compared
// with 32 real rand() calls (which are the next values of the stream and go into the FIFO like any other draw)
bool bulk_rand_selfcheck(std::vector<int32_t>& out)
{
    static char dummy[128];
    char* old = initstate(1u, dummy, sizeof(dummy));
    if (!old) return false;
    int32_t copy[32];
    std::memcpy(copy, old, sizeof(copy));
    setstate(old);
    const int type = copy[0] % 5, rear0 = copy[0] / 5;
    bool ok = type == 3 && rear0 >= 0 && rear0 < 31;
    uint32_t* state = reinterpret_cast<uint32_t*>(copy + 1);
    int r = rear0, f = (rear0 + 3) % 31;
    for (int i = 0; i < 32; ++i) {
        const int real = std::rand();
        out.push_back(real);
        if (ok) {
            const uint32_t val = state[f] += state[r];
            ok = (int32_t)(val >> 1) == real;
            if (++f == 31) f = 0;
            if (++r == 31) r = 0;
        }
    }
    return ok;
}

namespace {
void* pinned_alloc(size_t bytes)
{
    void* p = nullptr;
    if (lgca_b200_host_alloc(bytes, &p) != 0) {
        printf("ERROR in B200_Lattice::allocate_memory(): %s\n", lgca_b200_last_error());
        fflush(stdout);
        abort();
    }
    return p;
}
void pinned_free(void* p) { lgca_b200_host_free(p); }
} // namespace

template <Model model_>
void B200_Lattice<model_>::fail(const char* where, int rc)
{
    printf("ERROR in B200_Lattice::%s(): %s (code %d)\n", where, lgca_b200_last_error(), rc);
    fflush(stdout);
    abort();
}

template <Model model_>
B200_Lattice<model_>::B200_Lattice(const string test_case, const Real Re, const Real Ma_s, const int cg, const B200Options& opt)
    : Lattice<model_>(test_case, Re, Ma_s, cg), m_opt(opt)
{
    allocate_memory();
    this->m_rnd_cpu.fill_random(); // right after allocation, before any initialiser: rand() stream order
    create_device_lattice();
}

template <Model model_>
B200_Lattice<model_>::B200_Lattice(const string test_case, unsigned int dim_x, unsigned int dim_y, const int cg, char bf_dir,
                                   const B200Options& opt)
    : Lattice<model_>(test_case, dim_x, dim_y, cg, bf_dir), m_opt(opt)
{
    allocate_memory();
    this->m_rnd_cpu.fill_random();
    create_device_lattice();
}

template <Model model_>
B200_Lattice<model_>::~B200_Lattice()
{
    join_prefetch();
    if (m_h) lgca_b200_group_destroy(m_h);
    free_memory();
}

template <Model model_>
void B200_Lattice<model_>::allocate_memory()
{
    const size_t n = this->m_num_cells, nc = std::max<size_t>(this->m_num_coarse_cells, 1);
    this->m_cell_type_cpu = static_cast<CellType*>(pinned_alloc(n * sizeof(CellType)));
    std::memset(this->m_cell_type_cpu, 0, n * sizeof(CellType));
    if (m_opt.cell_fields) {
        this->m_cell_density_cpu  = static_cast<Real*>(pinned_alloc(n * sizeof(Real)));
        this->m_cell_momentum_cpu = static_cast<Real*>(pinned_alloc(2 * n * sizeof(Real)));
        std::memset(this->m_cell_density_cpu, 0, n * sizeof(Real));
        std::memset(this->m_cell_momentum_cpu, 0, 2 * n * sizeof(Real));
    }
    this->m_mean_density_cpu  = static_cast<Real*>(pinned_alloc(nc * sizeof(Real)));
    this->m_mean_momentum_cpu = static_cast<Real*>(pinned_alloc(2 * nc * sizeof(Real)));
    std::memset(this->m_mean_density_cpu, 0, nc * sizeof(Real));
    std::memset(this->m_mean_momentum_cpu, 0, 2 * nc * sizeof(Real));
    this->m_node_state_cpu.set_allocator(pinned_alloc, pinned_free);
    this->m_node_state_cpu.resize(n * 8);
    // the snapshot lives on the device; the host-side out buffer of the reference is not needed
    this->m_rnd_cpu.resize(n);
}

template <Model model_>
void B200_Lattice<model_>::free_memory()
{
    pinned_free(this->m_cell_type_cpu);
    pinned_free(this->m_cell_density_cpu);
    pinned_free(this->m_cell_momentum_cpu);
    pinned_free(this->m_mean_density_cpu);
    pinned_free(this->m_mean_momentum_cpu);
    this->m_cell_type_cpu = nullptr;
    this->m_cell_density_cpu = this->m_cell_momentum_cpu = this->m_mean_density_cpu = this->m_mean_momentum_cpu = nullptr;
}

template <Model model_>
void B200_Lattice<model_>::create_device_lattice()
{
    lgca_b200_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.model     = ModelDescriptor<model_>::C_ABI_ID;
    cfg.dim_x     = this->m_dim_x;
    cfg.dim_y     = this->m_dim_y;
    cfg.cg_radius = this->m_coarse_graining_radius;
    // the single-collision demo lattice (21 x 10, cg 1) violates dim_x % 2cg == 0 like in the reference
    if (cfg.cg_radius && (cfg.dim_x % (2 * cfg.cg_radius) || cfg.dim_y % (2 * cfg.cg_radius) || cfg.dim_x < 4 * cfg.cg_radius))
        cfg.cg_radius = 0;
    cfg.bf_dir = this->m_bf_dir;
    cfg.device = m_opt.device;
    cfg.k_fuse = m_opt.k_fuse;
    cfg.flags  = m_opt.cell_fields ? 0u : (uint32_t)LGCA_B200_FLAG_NO_CELL_FIELDS;
    if (m_opt.n_gpus < 1) m_opt.n_gpus = 1;
    std::vector<int> devs = m_opt.devices;
    if (devs.empty()) for (int i = 0; i < m_opt.n_gpus; ++i) devs.push_back(m_opt.device + i);
    if ((int)devs.size() != m_opt.n_gpus) {
        printf("ERROR in B200_Lattice::B200_Lattice(): %d device ordinals given for %d GPUs.\n", (int)devs.size(), m_opt.n_gpus);
        fflush(stdout);
        abort();
    }
    const int rc = lgca_b200_group_create(&cfg, m_opt.n_gpus, devs.data(), &m_h);
    if (rc) fail("B200_Lattice", rc);
}

template <Model model_>
void B200_Lattice<model_>::setup_parallel()
{
    lgca_b200_info info;
    const int rc = lgca_b200_group_get_info(m_h, &info);
    if (rc) fail("setup_parallel", rc);
    printf("B200 configuration parameters: %d GPU(s)%s, %u bit-planes of %u x %u words, %d fused steps per pass, "
           "%.1f MB on the device(s).\n\n", m_opt.n_gpus, m_opt.n_gpus > 1 ? " (row strips, peer-store halo ring)" : "",
           info.num_planes, info.y_rows, info.words_per_row, info.k_fuse, info.device_bytes / 1.0e6);
}

template <Model model_>
void B200_Lattice<model_>::copy_data_to_device()
{
    const int rc = lgca_b200_group_upload(m_h, this->m_node_state_cpu.ptr(), reinterpret_cast<const int32_t*>(this->m_cell_type_cpu),
                                    this->m_rnd_cpu.ptr());
    if (rc) fail("copy_data_to_device", rc);
    m_on_device = true;
}

template <Model model_>
void B200_Lattice<model_>::ensure_on_device()
{
    if (!m_on_device) copy_data_to_device();
}

template <Model model_>
void B200_Lattice<model_>::copy_data_from_device()
{
    ensure_on_device();
    const int rc = lgca_b200_group_download(m_h, this->m_node_state_cpu.ptr());
    if (rc) fail("copy_data_from_device", rc);
}

template <Model model_>
void B200_Lattice<model_>::collide_and_propagate(const bool /*p: ignored, as in the reference's live backend*/)
{
    collide_and_propagate_n(1);
}

template <Model model_>
void B200_Lattice<model_>::collide_and_propagate_n(int n_steps)
{
    ensure_on_device();
    const int rc = lgca_b200_group_step(m_h, n_steps);
    if (rc) fail("collide_and_propagate", rc);
}

template <Model model_>
void B200_Lattice<model_>::copy_data_to_output_buffer()
{
    ensure_on_device();
    const int rc = lgca_b200_group_snapshot(m_h);
    if (rc) fail("copy_data_to_output_buffer", rc);
}

template <Model model_>
void B200_Lattice<model_>::post_process()
{
    ensure_on_device();
    const bool coarse = this->m_num_coarse_cells > 0 && this->m_dim_x % (2 * this->m_coarse_graining_radius) == 0 &&
                        this->m_dim_x >= 4 * this->m_coarse_graining_radius;
    const bool lazy = m_opt.lazy_cell_fields && this->m_cell_density_cpu;
    const int rc = lgca_b200_group_post_process(m_h, lazy ? nullptr : this->m_cell_density_cpu, lazy ? nullptr : this->m_cell_momentum_cpu,
                                          coarse ? this->m_mean_density_cpu : nullptr,
                                          coarse ? this->m_mean_momentum_cpu : nullptr, m_opt.exact_post ? 1 : 0);
    if (rc) fail("post_process", rc);
    m_cell_fields_stale = lazy;
}

template <Model model_>
void B200_Lattice<model_>::sync_cell_fields()
{
    if (!m_cell_fields_stale) return;
    // the output buffer is unchanged since the post_process() that skipped them (a new snapshot is only taken by
    // copy_data_to_output_buffer, which the tick schedule always follows with post_process)
    const int rc = lgca_b200_group_post_process(m_h, this->m_cell_density_cpu, this->m_cell_momentum_cpu, nullptr, nullptr, 0);
    if (rc) fail("sync_cell_fields", rc);
    m_cell_fields_stale = false;
}

template <Model model_>
std::vector<Real> B200_Lattice<model_>::get_mean_velocity()
{
    std::vector<Real> mean_velocity(this->SPATIAL_DIM, 0.0);
    ensure_on_device();
    join_prefetch();
    if (m_opt.prefetch_draws && m_last_consumed) {
        // the tick's next call is apply_body_force(): draw what it is expected to consume while the GPU and the walk
        // below are busy (same FIFO, same order; joined before the FIFO is read)
        const size_t target = m_last_consumed + m_last_consumed / 4;
        if (draws_pending() < target) m_prefetch = std::thread([this, target] { draw_until(target); });
    }
    float out[2];
    if (this->m_num_cells / (size_t)m_opt.n_gpus <= ((size_t)1 << 28)) {
        // the reference's loop at one thread (src/omp_lattice.cpp:508-557): sequential float32 sums over the cells of
        // the output buffer -- the only order that reproduces its digits and hence the forcing decisions of the pipe /
        // Karman schedule.  The device reduces 1024-cell segments to integer summaries per float32 binade, the library
        // walks them in order (csrc/lgca_mv.cu): bit-equal, without a pass over the per-cell host fields.
        const int rc = lgca_b200_group_mean_velocity_exact(m_h, out);
        if (rc) fail("get_mean_velocity", rc);
        mean_velocity[0] = out[0];
        mean_velocity[1] = out[1];
        return mean_velocity;
    }
    // huge lattices (float32 sums saturate anyway): device reduction over the snapshot, double accumulation
    const int rc = lgca_b200_group_mean_velocity(m_h, out);
    if (rc) fail("get_mean_velocity", rc);
    mean_velocity[0] = out[0];
    mean_velocity[1] = out[1];
    return mean_velocity;
}

template <Model model_>
void B200_Lattice<model_>::apply_body_force(const int forcing)
{
    ensure_on_device();
    join_prefetch();
    // The reference draws `rand() % num_cells` one at a time until `forcing` particles are reverted or
    // 2*num_cells draws are spent (do-while: at least one draw; src/omp_lattice.cpp:254-346).  Here rand() values
    // are drawn ahead into a FIFO, handed to the device in order, and the unconsumed ones are kept for the next
    // call, so the stream position after the call equals the reference's as seen by this lattice.
    const size_t it_max = 2 * this->m_num_cells;
    size_t it = 0;
    long   remaining = (long)(unsigned int)forcing; // `unsigned int < int` compares as unsigned in the reference
    bool   first = true;
    if (m_draw_head > (1u << 20) && m_draw_head * 2 > m_draws.size()) { // drop the consumed prefix now and then
        m_draws.erase(m_draws.begin(), m_draws.begin() + m_draw_head);
        m_draw_head = 0;
    }
    while ((first || remaining > 0) && it < it_max) {
        size_t want = (size_t)std::max<double>(256.0, (double)std::max<long>(remaining, 1) * m_draws_per_hit * 1.25);
        want = std::min(want, it_max - it);
#ifdef LGCA_BF_TRACE
        const auto tr0 = std::chrono::steady_clock::now();
        const size_t had = draws_pending();
#endif
        draw_until(want);
#ifdef LGCA_BF_TRACE
        const auto tr1 = std::chrono::steady_clock::now();
#endif
        size_t consumed = 0;
        uint32_t reverted = 0;
        const int rc = lgca_b200_group_body_force(m_h, (int)remaining, m_draws.data() + m_draw_head, want, &consumed, &reverted);
        if (rc) fail("apply_body_force", rc);
#ifdef LGCA_BF_TRACE
        const auto tr2 = std::chrono::steady_clock::now();
        fprintf(stderr, "bf: remaining %ld want %zu had %zu consumed %zu reverted %u dph %.2f draw %.3f ms device %.3f ms\n", remaining, want, had,
                consumed, reverted, m_draws_per_hit, std::chrono::duration<double, std::milli>(tr1 - tr0).count(),
                std::chrono::duration<double, std::milli>(tr2 - tr1).count());
#endif
        m_draw_head += consumed;
        it += consumed;
        remaining -= reverted;
        if (reverted > 0) m_draws_per_hit = 0.5 * m_draws_per_hit + 0.5 * std::min(1.0e4, (double)consumed / reverted);
        else m_draws_per_hit = std::min(1.0e4, m_draws_per_hit * 2.0);
        first = false;
        if (consumed == 0) break;
    }
    m_last_consumed = it;
}

template <Model model_>
void B200_Lattice<model_>::draw_until(size_t pending)
{
    static int bulk_state = 0; // 0 = unchecked, 1 = usable, -1 = not this libc's generator: plain rand()
    if (m_opt.bulk_rand && bulk_state == 0 && draws_pending() < pending) bulk_state = bulk_rand_selfcheck(m_draws) ? 1 : -1;
    if (m_opt.bulk_rand && bulk_state == 1 && draws_pending() < pending) {
        if (bulk_rand_fill(m_draws, pending - draws_pending())) return;
        bulk_state = -1;
    }
    while (draws_pending() < pending) m_draws.push_back(std::rand());
}

template <Model model_>
unsigned long B200_Lattice<model_>::get_n_particles()
{
    if (!m_on_device) return Lattice<model_>::get_n_particles(); // still being set up on the host
    uint64_t n = 0;
    const int rc = lgca_b200_group_count_particles(m_h, &n);
    if (rc) fail("get_n_particles", rc);
    this->m_num_particles = n;
    return n;
}

template <Model model_>
void B200_Lattice<model_>::synchronize()
{
    const int rc = lgca_b200_group_sync(m_h);
    if (rc) fail("synchronize", rc);
}

template <Model model_>
double B200_Lattice<model_>::timed_steps(int n_steps)
{
    ensure_on_device();
    float ms = 0;
    const int rc = lgca_b200_group_timed_steps(m_h, n_steps, &ms);
    if (rc) fail("timed_steps", rc);
    return ms;
}

} // namespace lgca

// test hooks (tests/test_bulk_rand.py, no GPU needed): n values of the rand() stream drawn in bulk / the self-check
extern "C" int lgca_host_bulk_rand(int32_t* out, size_t n)
{
    std::vector<int32_t> v;
    if (!lgca::bulk_rand_fill(v, n)) return 0;
    std::memcpy(out, v.data(), n * sizeof(int32_t));
    return 1;
}
extern "C" int lgca_host_bulk_rand_selfcheck(int32_t* out32)
{
    std::vector<int32_t> v;
    const bool ok = lgca::bulk_rand_selfcheck(v);
    std::memcpy(out32, v.data(), 32 * sizeof(int32_t));
    return ok ? 1 : 0;
}

namespace lgca {

template class B200_Lattice<Model::HPP>;
template class B200_Lattice<Model::FHP_I>;
template class B200_Lattice<Model::FHP_II>;
template class B200_Lattice<Model::FHP_III>;

} // namespace lgca
