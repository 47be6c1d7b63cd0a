// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// C-ABI driver around the UNMODIFIED reference CPU backend (OMP_Lattice<Model>), compiled in
// place from /root/reference/src/{lattice,omp_lattice}.cpp by oracle/Makefile into
// oracle/_ref/liblgca_ref.so.  Nothing from the reference is copied into this repository: this
// file only *calls* the reference's public/protected interface
//   (reference: src/lattice.h:32-238, src/omp_lattice.h:29-78).
//
// Used by (a) tests/ to pin oracle/lgca_oracle.c and the CUDA path against the real reference and
// (b) bench.py's `--impl reference` / `cpu_baseline` legs to time the reference's own hot loop
// (src/omp_lattice.cpp:100-249) on the host cores.
//
// System headers first, then open up protected/private so the driver can reach the raw arrays
// (state bytes, cell types, chirality bits) that parity is defined on.

#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <fstream>
#include <cmath>
#include <vector>
#include <string>
#include <cassert>
#include <chrono>
#include <limits>
#include <cstdint>
#include <omp.h>
#include <unistd.h>
#include <fcntl.h>
#include <cstring>
#include "tclap/CmdLine.h"   // third-party header the reference pulls in; include it before the access hack

#define protected public
#define private public
#include "omp_lattice.h"
#undef protected
#undef private

namespace {

using lgca::Model;
using lgca::CellType;
using lgca::Real;

struct IRef {
    virtual ~IRef() {}
    virtual unsigned dim_x() = 0;
    virtual unsigned dim_y() = 0;
    virtual unsigned coarse_dim_x() = 0;
    virtual unsigned coarse_dim_y() = 0;
    virtual int num_dir() = 0;
    virtual float u() = 0;
    virtual void apply_bc(const char* name) = 0;
    virtual void init(const char* name) = 0;
    virtual uint8_t* state() = 0;
    virtual uint8_t* state_out() = 0;
    virtual int32_t* cell_type() = 0;
    virtual uint8_t* rnd() = 0;
    virtual void step(int n) = 0;
    virtual void body_force(int forcing) = 0;
    virtual void snapshot() = 0;
    virtual void post_process() = 0;
    virtual void mean_velocity(float* out) = 0;
    virtual unsigned long n_particles() = 0;
    virtual float* cell_density() = 0;
    virtual float* cell_momentum() = 0;
    virtual float* mean_density() = 0;
    virtual float* mean_momentum() = 0;
    virtual size_t initial_forcing() = 0;
    virtual size_t equilibrium_forcing() = 0;
    virtual void set_bf_dir(char c) = 0;
    virtual void resize(unsigned dx, unsigned dy) = 0;
};

template <Model M>
struct RefImpl : IRef {
    lgca::OMP_Lattice<M>* lat;
    RefImpl(const char* tc, float Re, float Ma, int cg) { lat = new lgca::OMP_Lattice<M>(tc, Re, Ma, cg); }
    ~RefImpl() override { delete lat; }
    unsigned dim_x() override { return lat->dim_x(); }
    unsigned dim_y() override { return lat->dim_y(); }
    unsigned coarse_dim_x() override { return lat->coarse_dim_x(); }
    unsigned coarse_dim_y() override { return lat->coarse_dim_y(); }
    int num_dir() override { return (int)lgca::ModelDescriptor<M>::NUM_DIR; }
    float u() override { return lat->u(); }
    void apply_bc(const char* name) override {
        std::string s(name);
        if      (s == "periodic")           lat->apply_bc_periodic();
        else if (s == "pipe")               lat->apply_bc_pipe();
        else if (s == "karman")             lat->apply_bc_karman_vortex_street();
        else if (s == "reflecting_back")    lat->apply_bc_reflecting("back");
        else if (s == "reflecting_forward") lat->apply_bc_reflecting("forward");
        else { fprintf(stderr, "ref_driver: unknown bc %s\n", name); abort(); }
    }
    void init(const char* name) override {
        std::string s(name);
        if      (s == "random")           lat->init_random();
        else if (s == "diffusion")        lat->init_diffusion();
        else if (s == "single_collision") lat->init_single_collision();
        else if (s == "zero")             lat->init_zero();
        else { fprintf(stderr, "ref_driver: unknown init %s\n", name); abort(); }
    }
    uint8_t* state() override { return lat->m_node_state_cpu.ptr(); }
    uint8_t* state_out() override { return lat->m_node_state_out_cpu.ptr(); }
    int32_t* cell_type() override { return reinterpret_cast<int32_t*>(lat->m_cell_type_cpu); }
    uint8_t* rnd() override { return lat->m_rnd_cpu.ptr(); }
    void step(int n) override { for (int i = 0; i < n; ++i) lat->collide_and_propagate(false); }
    void body_force(int forcing) override { lat->apply_body_force(forcing); }
    void snapshot() override { lat->copy_data_to_output_buffer(); }
    void post_process() override { lat->post_process(); }
    void mean_velocity(float* out) override {
        std::vector<Real> v = lat->get_mean_velocity();
        out[0] = v[0]; out[1] = v[1];
    }
    unsigned long n_particles() override { return lat->get_n_particles(); }
    float* cell_density() override { return lat->m_cell_density_cpu; }
    float* cell_momentum() override { return lat->m_cell_momentum_cpu; }
    float* mean_density() override { return lat->m_mean_density_cpu; }
    float* mean_momentum() override { return lat->m_mean_momentum_cpu; }
    size_t initial_forcing() override { return lat->get_initial_forcing(); }
    size_t equilibrium_forcing() override { return lat->get_equilibrium_forcing(); }
    void set_bf_dir(char c) override { lat->m_bf_dir = c; }
    // Test-only: reach lattice shapes the reference ctor cannot produce (e.g. non-square boxes)
    // by overriding the dims and re-running the reference's own allocation + table set-up.
    void resize(unsigned dx, unsigned dy) override {
        lat->free_memory();
        lat->m_dim_x = dx;
        lat->m_dim_y = dy;
        lat->m_num_cells = (size_t)dx * dy;
        lat->m_num_nodes = lat->m_num_cells * lat->NUM_DIR;
        unsigned r = lat->m_coarse_graining_radius;
        lat->m_coarse_dim_x = dx / (2 * r);
        lat->m_coarse_dim_y = dy / (2 * r);
        lat->m_num_coarse_cells = (size_t)lat->m_coarse_dim_x * lat->m_coarse_dim_y;
        lat->allocate_memory();
        lat->m_rnd_cpu.fill_random();
        delete lat->m_model;
        lat->m_model = new lgca::ModelDescriptor<M>(dx, dy);
    }
};

// The reference ctor prints a parameter banner to stdout; keep test logs quiet.
struct QuietStdout {
    int saved;
    QuietStdout() {
        fflush(stdout);
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 1); close(nul); }
    }
    ~QuietStdout() {
        fflush(stdout);
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};

} // namespace

extern "C" {

void* lgca_ref_create(int model, const char* test_case, float Re, float Ma, int cg) {
    QuietStdout q;
    switch (model) {
    case 0: return new RefImpl<Model::HPP>(test_case, Re, Ma, cg);
    case 1: return new RefImpl<Model::FHP_I>(test_case, Re, Ma, cg);
    case 2: return new RefImpl<Model::FHP_II>(test_case, Re, Ma, cg);
    case 3: return new RefImpl<Model::FHP_III>(test_case, Re, Ma, cg);
    }
    return nullptr;
}
void lgca_ref_destroy(void* h) { delete static_cast<IRef*>(h); }
void lgca_ref_resize(void* h, unsigned dx, unsigned dy) { static_cast<IRef*>(h)->resize(dx, dy); }
unsigned lgca_ref_dim_x(void* h) { return static_cast<IRef*>(h)->dim_x(); }
unsigned lgca_ref_dim_y(void* h) { return static_cast<IRef*>(h)->dim_y(); }
unsigned lgca_ref_coarse_dim_x(void* h) { return static_cast<IRef*>(h)->coarse_dim_x(); }
unsigned lgca_ref_coarse_dim_y(void* h) { return static_cast<IRef*>(h)->coarse_dim_y(); }
int lgca_ref_num_dir(void* h) { return static_cast<IRef*>(h)->num_dir(); }
float lgca_ref_u(void* h) { return static_cast<IRef*>(h)->u(); }
void lgca_ref_apply_bc(void* h, const char* name) { static_cast<IRef*>(h)->apply_bc(name); }
void lgca_ref_init(void* h, const char* name) { static_cast<IRef*>(h)->init(name); }
uint8_t* lgca_ref_state(void* h) { return static_cast<IRef*>(h)->state(); }
uint8_t* lgca_ref_state_out(void* h) { return static_cast<IRef*>(h)->state_out(); }
int32_t* lgca_ref_cell_type(void* h) { return static_cast<IRef*>(h)->cell_type(); }
uint8_t* lgca_ref_rnd(void* h) { return static_cast<IRef*>(h)->rnd(); }
void lgca_ref_step(void* h, int n) { static_cast<IRef*>(h)->step(n); }
void lgca_ref_body_force(void* h, int forcing) { static_cast<IRef*>(h)->body_force(forcing); }
void lgca_ref_snapshot(void* h) { static_cast<IRef*>(h)->snapshot(); }
void lgca_ref_post_process(void* h) { static_cast<IRef*>(h)->post_process(); }
void lgca_ref_mean_velocity(void* h, float* out) { static_cast<IRef*>(h)->mean_velocity(out); }
unsigned long lgca_ref_n_particles(void* h) { return static_cast<IRef*>(h)->n_particles(); }
float* lgca_ref_cell_density(void* h) { return static_cast<IRef*>(h)->cell_density(); }
float* lgca_ref_cell_momentum(void* h) { return static_cast<IRef*>(h)->cell_momentum(); }
float* lgca_ref_mean_density(void* h) { return static_cast<IRef*>(h)->mean_density(); }
float* lgca_ref_mean_momentum(void* h) { return static_cast<IRef*>(h)->mean_momentum(); }
size_t lgca_ref_initial_forcing(void* h) { return static_cast<IRef*>(h)->initial_forcing(); }
size_t lgca_ref_equilibrium_forcing(void* h) { return static_cast<IRef*>(h)->equilibrium_forcing(); }
void lgca_ref_set_bf_dir(void* h, char c) { static_cast<IRef*>(h)->set_bf_dir(c); }
int lgca_ref_max_threads(void) { return omp_get_max_threads(); }
void lgca_ref_set_threads(int n) { omp_set_num_threads(n); }
void lgca_ref_srand(unsigned seed) { srand(seed); }
int lgca_ref_rand(void) { return rand(); }

} // extern "C"
