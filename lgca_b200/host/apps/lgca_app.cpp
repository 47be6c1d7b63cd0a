// Headless LGCA apps on the B200 backend: lgca-pipe, lgca-karman, lgca-diffusion, lgca-single, lgca-box,
// lgca-periodic (one source, the test case is fixed per binary with -DLGCA_APP_CASE="...").
//
// They replace the reference's Qt/VTK viewers (apps/{pipe,karman,diffusion,single}/*_viewer.cpp) and its stale CLI
// (apps/{box,periodic}/main.cpp) with the same set-up sequence and the canonical, serialised tick schedule of the
// viewers (apps/pipe/pipe_viewer.cpp:100-184):
//     mean velocity -> (body force) -> PP_INTERVAL x collide_and_propagate -> snapshot -> post_process [-> write]
// Flags follow src/utils.h:49-83 (-r/--Re, -m/--Ma, -d/--n-dir, -s/--steps, -c/--cg-radius, -w/--write-steps,
// --device, -o/--output) plus --gpus / --devices (row strips over several GPUs), --model, --pp-interval, --dims, --k-fuse, --hash-every, --no-cell-fields, --quiet.
#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <string>

#include "../b200_lattice.h"
#include "../lgca_io_png.h"
#include "../lgca_io_vti.h"

#ifndef LGCA_APP_CASE
#define LGCA_APP_CASE "pipe"
#endif

using namespace lgca;

namespace {

struct Args {
    std::string test_case = LGCA_APP_CASE;
    std::string model;
    Real        Re = 80.0, Ma = 0.3;
    int         steps = 50, cg = 10, write_steps = 0, pp_interval = 5, device = 0, k_fuse = 0, gpus = 1;
    std::vector<int> devices;
    int         hash_every = 0;
    unsigned    dim_x = 0, dim_y = 0;
    bool        cell_fields = true, quiet = false, bc_forward = false, plain_rand = false;
    std::string output = "none", out_dir = "./", scalars = "Mean momentum";
    int         zoom = 1;
};

uint64_t fnv1a64(const uint8_t* p, size_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

void usage(const char* argv0)
{
    printf("usage: %s [-r Re] [-m Ma] [-d 4|6|7] [--model HPP|FHP_I|FHP_II|FHP_III] [-s steps] [-c cg-radius]\n"
           "          [-w write-steps] [--pp-interval n] [--dims X Y] [--device n] [--gpus n | --devices a,b,..] [--k-fuse k]\n"
           "          [-o none|vti|png] [--out-dir d] [--scalars \"Mean momentum\"] [--zoom n]\n"
           "          [--hash-every n] [--bounce forward|back] [--no-cell-fields] [--quiet] [--plain-rand]\n", argv0);
}

bool parse(int argc, char** argv, Args& a)
{
    // per-app defaults = the compile-time constants of the reference's viewers
    if (a.test_case == "pipe")      { a.Re = 80;  a.Ma = 0.3; a.cg = 10; a.pp_interval = 5; a.model = "FHP_III"; }
    if (a.test_case == "karman")    { a.Re = 80;  a.Ma = 0.3; a.cg = 20; a.pp_interval = 5; a.model = "FHP_III"; }
    if (a.test_case == "diffusion") { a.Re = 200; a.Ma = 0.2; a.cg = 1;  a.pp_interval = 1; a.model = "FHP_III"; }
    if (a.test_case == "collision") { a.Re = 80;  a.Ma = 0.2; a.cg = 1;  a.pp_interval = 1; a.model = "FHP_III"; }
    if (a.test_case == "box" || a.test_case == "periodic") { a.Re = 255; a.Ma = 0.2; a.cg = 16; a.pp_interval = 10; a.model = "FHP_III"; }
    for (int i = 1; i < argc; ++i) {
        std::string f = argv[i];
        auto next = [&](const char* what) -> const char* {
            if (i + 1 >= argc) { printf("ERROR in main(): missing value for %s\n", what); exit(2); }
            return argv[++i];
        };
        if (f == "-r" || f == "--Re") a.Re = (Real)atof(next("Re"));
        else if (f == "-m" || f == "--Ma") a.Ma = (Real)atof(next("Ma"));
        else if (f == "-d" || f == "--n-dir") { int d = atoi(next("n-dir")); a.model = d == 4 ? "HPP" : (d == 6 ? "FHP_I" : "FHP_III"); }
        else if (f == "--model") a.model = next("model");
        else if (f == "-s" || f == "--steps") a.steps = atoi(next("steps"));
        else if (f == "-c" || f == "--cg-radius") a.cg = atoi(next("cg-radius"));
        else if (f == "-w" || f == "--write-steps") a.write_steps = atoi(next("write-steps"));
        else if (f == "--pp-interval") a.pp_interval = atoi(next("pp-interval"));
        else if (f == "--dims") { a.dim_x = (unsigned)atoi(next("dims")); a.dim_y = (unsigned)atoi(next("dims")); }
        else if (f == "--device") a.device = atoi(next("device"));
        else if (f == "--gpus") a.gpus = atoi(next("gpus"));
        else if (f == "--devices") { // explicit ordinals of the strips, e.g. 0,1,2,3 (a device may repeat: testing)
            std::string v = next("devices");
            a.devices.clear();
            for (size_t p0 = 0; p0 <= v.size();) {
                size_t p1 = v.find(',', p0);
                if (p1 == std::string::npos) p1 = v.size();
                if (p1 > p0) a.devices.push_back(atoi(v.substr(p0, p1 - p0).c_str()));
                p0 = p1 + 1;
            }
            a.gpus = (int)a.devices.size();
        }
        else if (f == "--k-fuse") a.k_fuse = atoi(next("k-fuse"));
        else if (f == "-o" || f == "--output") a.output = next("output");
        else if (f == "--out-dir") a.out_dir = next("out-dir");
        else if (f == "--scalars") a.scalars = next("scalars");   // "Cell density" | "Cell momentum" | "Mean density" | "Mean momentum"
        else if (f == "--zoom") a.zoom = atoi(next("zoom"));
        else if (f == "--hash-every") a.hash_every = atoi(next("hash-every"));
        else if (f == "--bounce") a.bc_forward = std::string(next("bounce")) == "forward";
        else if (f == "--no-cell-fields") a.cell_fields = false;
        else if (f == "--quiet") a.quiet = true;
        else if (f == "--plain-rand") a.plain_rand = true;   // body-force draws through rand() calls only (A-B of bulk_rand)
        else if (f == "-p" || f == "--parallel") { std::string p = next("parallel"); if (p != "B200" && p != "CUDA") { printf("ERROR in main(): Invalid parallelization type %s (only B200).\n", p.c_str()); exit(2); } }
        else if (f == "--bf-steps" || f == "--bf-int" || f == "--blocksize") next(f.c_str()); // accepted, unused (as in the reference's live apps)
        else if (f == "-h" || f == "--help") { usage(argv[0]); exit(0); }
        else { printf("ERROR in main(): unknown flag %s\n", f.c_str()); usage(argv[0]); return false; }
    }
    return true;
}

template <Model M>
int run(const Args& a)
{
    B200Options opt;
    opt.device = a.device;
    opt.n_gpus = a.gpus;
    opt.devices = a.devices;
    opt.k_fuse = a.k_fuse;
    opt.cell_fields = a.cell_fields;
    opt.lazy_cell_fields = true; // the per-cell fields cross PCIe only when a writer needs them
    opt.prefetch_draws = true;   // nothing else in this process calls rand() during the tick loop
    opt.bulk_rand = !a.plain_rand;
    // --quiet: the base-class ctor prints the parameter banner; silence fd 1 around the construction only
    int saved_fd = -1;
    if (a.quiet) {
        fflush(stdout);
        saved_fd = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 1); close(nul); }
    }
    B200_Lattice<M>* lat = a.dim_x ? new B200_Lattice<M>(a.test_case, a.dim_x, a.dim_y, a.cg,
                                                         (a.test_case == "pipe" || a.test_case == "karman") ? 'x' : 0, opt)
                                   : new B200_Lattice<M>(a.test_case, a.Re, a.Ma, a.cg, opt);
    if (saved_fd >= 0) { fflush(stdout); dup2(saved_fd, 1); close(saved_fd); }

    // boundary conditions + initial particles: apps/periodic/main.cpp:100-148 and the viewer ctors
    const std::string& tc = a.test_case;
    if (tc == "pipe" || tc == "collision") lat->apply_bc_pipe();
    else if (tc == "karman") lat->apply_bc_karman_vortex_street();
    else if (tc == "box" || tc == "diffusion") lat->apply_bc_reflecting(a.bc_forward ? "forward" : "back");
    else lat->apply_bc_periodic();
    if (tc == "collision") lat->init_single_collision();
    else if (tc == "diffusion") lat->init_diffusion();
    else lat->init_random();
    const unsigned long particles_start = lat->get_n_particles();

    lat->copy_data_to_device();
    lat->copy_data_to_output_buffer();
    lat->post_process();
    int forcing = (int)lat->get_initial_forcing();
    lat->setup_parallel();
    IoVti<M> vti(lat);
    IoPng<M> png(lat, a.scalars, (unsigned)a.zoom);
    const bool forced = (tc == "pipe" || tc == "karman");

    auto print_hash = [&](size_t step) {
        lat->copy_data_from_device();
        printf("HASH step %zu %016llx\n", step, (unsigned long long)fnv1a64(lat->state_bytes(), lat->num_cells()));
    };
    if (a.hash_every) print_hash(0);

    size_t steps = 0, ticks = 0;
    double sim_seconds = 0, t_mv = 0, t_bf = 0, t_pp = 0;
    std::vector<Real> mv(2, 0.0);
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto loop_start = clk::now();
    while ((int)steps < a.steps) {
        if (forced) {
            auto t0 = clk::now();
            mv = lat->get_mean_velocity();
            auto t1 = clk::now();
            if (mv[0] < lat->u()) {
                if (mv[0] > 0.9 * lat->u()) forcing = (int)lat->get_equilibrium_forcing();
                lat->apply_body_force(forcing);
            }
            lat->synchronize();
            t_mv += secs(t0, t1);
            t_bf += secs(t1, clk::now());
        }
        const int n = std::min(a.pp_interval, a.steps - (int)steps);
        lat->synchronize();
        auto t0 = clk::now();
        lat->collide_and_propagate_n(n);
        lat->synchronize();
        auto t1 = clk::now();
        sim_seconds += secs(t0, t1);
        steps += n;
        ++ticks;
        lat->copy_data_to_output_buffer();
        lat->post_process();
        t_pp += secs(t1, clk::now());
        if (a.write_steps > 0 && steps % a.write_steps == 0) {
            if (!a.quiet) printf("Executing step %zu... mean velocity (%6.4f, %6.4f)\n", steps, mv[0], mv[1]);
            if (a.output == "vti") vti.write(steps, a.out_dir);
            else if (a.output == "png") png.write(steps, a.out_dir);
        }
        if (a.hash_every && steps % a.hash_every == 0) print_hash(steps);
    }
    const double loop_seconds = secs(loop_start, clk::now());
    const unsigned long particles_end = lat->get_n_particles();
    if (particles_end == particles_start) printf("Error check PASSED: There is no difference in the number of particles.\n");
    else printf("Error check FAILED: There is a difference in the number of particles of %ld.\n", (long)particles_end - (long)particles_start);
    printf("Total simulation time: %e s for %zu simulation steps.\n", sim_seconds, steps);
    if (sim_seconds > 0) printf("Average MNUPS: %.0f\n", (double)lat->num_cells() * steps / (sim_seconds * 1.0e06));
    // the whole tick loop as the viewers run it: mean velocity -> body force -> steps -> snapshot + post-process
    printf("Tick loop: %e s wall for %zu steps in %zu ticks (mean velocity %e s, body force %e s, stepping %e s, "
           "snapshot + post-process %e s); %.0f site updates/s end to end.\n", loop_seconds, steps, ticks, t_mv, t_bf, sim_seconds,
           t_pp, loop_seconds > 0 ? (double)lat->num_cells() * steps / loop_seconds : 0.0);
    if (forced) printf("Mean velocity: %.9g %.9g\n", mv[0], mv[1]);
    print_hash(steps);
    delete lat;
    return particles_end == particles_start ? 0 : 1;
}

} // namespace

int main(int argc, char** argv)
{
    Args a;
    if (!parse(argc, argv, a)) return 2;
    if (a.model == "HPP") return run<Model::HPP>(a);
    if (a.model == "FHP_I") return run<Model::FHP_I>(a);
    if (a.model == "FHP_II") return run<Model::FHP_II>(a);
    if (a.model == "FHP_III") return run<Model::FHP_III>(a);
    printf("ERROR in main(): Invalid model %s.\n", a.model.c_str());
    return 2;
}
