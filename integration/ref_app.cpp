// Drop-in proof: the REFERENCE's own Lattice<Model> base (its sizing, BC painters, initialisers, forcing formulas --
// compiled from /root/reference/src/lattice.cpp) with the B200 backend of integration/b200_lattice.h, on the canonical
// tick schedule of the viewers (apps/pipe/pipe_viewer.cpp:100-184).  Prints FNV-1a-64 hashes of the state bytes that
// tests/test_integration_ref.py compares with the reference's known answers.  Built by oracle/Makefile (target `ref`).
//     ref_b200_app <case: collision|pipe|karman|periodic> <model: 4|6|7> <Re> <Ma> <cg> <steps> <hash-every> [n_gpus]
#include <cstring>
#include <string>

#include "b200_lattice.h"

using namespace lgca;

static unsigned long long fnv(const unsigned char* p, size_t n)
{
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

template <Model M>
static int run(const std::string& tc, float Re, float Ma, int cg, int steps, int hash_every, int n_gpus)
{
    B200_Lattice<M>* lat = new B200_Lattice<M>(tc, Re, Ma, cg, n_gpus);
    if (tc == "pipe" || tc == "collision") lat->apply_bc_pipe();
    else if (tc == "karman") lat->apply_bc_karman_vortex_street();
    else lat->apply_bc_periodic();
    if (tc == "collision") lat->init_single_collision(); else lat->init_random();
    const unsigned long particles = lat->get_n_particles();
    lat->copy_data_to_device();
    lat->copy_data_to_output_buffer();
    lat->post_process();
    int forcing = (int)lat->get_initial_forcing();
    const bool forced = tc == "pipe" || tc == "karman";
    const int pp = tc == "collision" ? 1 : 5;
    printf("HASH step 0 %016llx\n", fnv(lat->state_bytes(), lat->num_cells()));
    for (int done = 0; done < steps;) {
        if (forced) {
            std::vector<Real> mv = lat->get_mean_velocity();
            if (mv[0] < lat->u()) {
                if (mv[0] > 0.9 * lat->u()) forcing = (int)lat->get_equilibrium_forcing();
                lat->apply_body_force(forcing);
            }
        }
        for (int s = 0; s < pp; ++s) lat->collide_and_propagate(false);   // one virtual call per step, like the viewers
        done += pp;
        lat->copy_data_to_output_buffer();
        lat->post_process();
        if (hash_every && done % hash_every == 0) {
            lat->copy_data_from_device();
            printf("HASH step %d %016llx\n", done, fnv(lat->state_bytes(), lat->num_cells()));
        }
    }
    lat->copy_data_from_device();
    const unsigned long end = lat->get_n_particles();   // base-class reader over the host mirror
    printf("PARTICLES %lu %lu\n", particles, end);
    delete lat;
    return particles == end ? 0 : 1;
}

int main(int argc, char** argv)
{
    if (argc < 8) { printf("usage: %s case n_dir Re Ma cg steps hash_every [n_gpus]\n", argv[0]); return 2; }
    const std::string tc = argv[1];
    const int nd = atoi(argv[2]), cg = atoi(argv[5]), steps = atoi(argv[6]), he = atoi(argv[7]), ng = argc > 8 ? atoi(argv[8]) : 1;
    const float Re = (float)atof(argv[3]), Ma = (float)atof(argv[4]);
    if (nd == 4) return run<Model::HPP>(tc, Re, Ma, cg, steps, he, ng);
    if (nd == 6) return run<Model::FHP_I>(tc, Re, Ma, cg, steps, he, ng);
    return run<Model::FHP_III>(tc, Re, Ma, cg, steps, he, ng);
}
