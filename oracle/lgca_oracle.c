/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the LGCA hot path (see lgca_oracle.h for the contract).
 *
 * Plain-C restatement of keva92/lgca's CPU path.  Every function cites the reference lines it
 * follows (paths relative to /root/reference).  Parity status: PINNED by tests/test_oracle.py
 * against SURVEY.md Appendix B known answers and against oracle/_ref (the unmodified reference).
 *
 * The restatement deliberately keeps the reference's per-cell, table-driven formulation (byte per
 * cell, offset tables with periodic corrections) so that it shares no structure with the bit-plane
 * CUDA kernels it is used to check.
 */
#define _GNU_SOURCE
#include "lgca_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * glibc rand(): TYPE_3 additive feedback generator x^31 + x^3 + 1 (glibc stdlib/random_r.c,
 * __srandom_r / __random_r; glibc 2.39).  The reference draws from it in src/lgca_bitset.h:223,
 * src/utils.h:119-122 and src/omp_lattice.cpp:269 and never seeds it (=> seed 1).
 * ---------------------------------------------------------------------------------------------- */
void lgca_oracle_srand(lgca_oracle_rng* g, unsigned seed)
{
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    g->r[0] = word;
    for (int i = 1; i < 31; ++i) {
        long hi = word / 127773;
        long lo = word % 127773;
        long w  = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word    = (int32_t)w;
        g->r[i] = word;
    }
    g->f = 3;
    g->b = 0;
    for (int i = 0; i < 310; ++i) (void)lgca_oracle_rand(g);
}

int lgca_oracle_rand(lgca_oracle_rng* g)
{
    uint32_t val = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
    g->r[g->f]   = (int32_t)val;
    int result   = (int)(val >> 1);
    if (++g->f >= 31) g->f = 0;
    if (++g->b >= 31) g->b = 0;
    return result;
}

/* ------------------------------------------------------------------------------------------------
 * Model tables: src/lgca_models.h:35-52 (HPP), :228-247 (FHP-I), :434-454 (FHP-II), :652-672 (FHP-III)
 * ---------------------------------------------------------------------------------------------- */
int lgca_oracle_num_dir(int model)
{
    switch (model) {
    case LGCA_OR_HPP:     return 4;
    case LGCA_OR_FHP_I:   return 6;
    case LGCA_OR_FHP_II:  return 7;
    case LGCA_OR_FHP_III: return 7;
    }
    return -1;
}

static const int HPP_INV[4]   = {2, 3, 0, 1};
static const int HPP_MIR_X[4] = {0, 3, 2, 1};
static const int HPP_MIR_Y[4] = {2, 1, 0, 3};
static const int FHP_INV[7]   = {3, 4, 5, 0, 1, 2, 6};
static const int FHP_MIR_X[7] = {0, 5, 4, 3, 2, 1, 6};
static const int FHP_MIR_Y[7] = {3, 2, 1, 0, 5, 4, 6};

static const int* inv_dir(int model)   { return model == LGCA_OR_HPP ? HPP_INV : FHP_INV; }
static const int* mir_dir_x(int model) { return model == LGCA_OR_HPP ? HPP_MIR_X : FHP_MIR_X; }
static const int* mir_dir_y(int model) { return model == LGCA_OR_HPP ? HPP_MIR_Y : FHP_MIR_Y; }

/* Lattice vectors as the reference's `Real` (float) constants, SIN = float(sin(M_PI/3)). */
static void lattice_vecs(int model, float vx[7], float vy[7])
{
    if (model == LGCA_OR_HPP) {
        const float x[4] = {1.0f, 0.0f, -1.0f, 0.0f};
        const float y[4] = {0.0f, 1.0f, 0.0f, -1.0f};
        for (int d = 0; d < 4; ++d) { vx[d] = x[d]; vy[d] = y[d]; }
        return;
    }
    const float s = (float)sin(M_PI / 3);
    const float x[7] = {1.0f, 0.5f, -0.5f, -1.0f, -0.5f, 0.5f, 0.0f};
    const float y[7] = {0.0f, s, s, 0.0f, -s, -s, 0.0f};
    for (int d = 0; d < 7; ++d) { vx[d] = x[d]; vy[d] = y[d]; }
}

/* ------------------------------------------------------------------------------------------------
 * Sizing: Lattice<M>::Lattice, src/lattice.cpp:29-157 (float/double mix reproduced statement by
 * statement: `Real` members are float, literals are double).
 * ---------------------------------------------------------------------------------------------- */
static void finish_dims(lgca_oracle_params* p, int cg_radius)
{
    /* src/lattice.cpp:144: computed in unsigned int arithmetic (overflows at 2^32 cells) */
    p->num_cells        = (uint32_t)(p->dim_x * p->dim_y);
    p->cg_radius        = (uint32_t)cg_radius;
    p->coarse_dim_x     = p->dim_x / (2u * p->cg_radius); /* :152 */
    p->coarse_dim_y     = p->dim_y / (2u * p->cg_radius); /* :153 */
    p->num_coarse_cells = (uint64_t)p->coarse_dim_x * p->coarse_dim_y;
}

static void physics(lgca_oracle_params* p, int model, float Re, float Ma_s)
{
    const float    rho = 1.0f, c = 1.0f;
    const unsigned SPATIAL_DIM = 2;
    p->model   = model;
    p->num_dir = lgca_oracle_num_dir(model);
    p->Re      = Re;
    p->Ma_s    = Ma_s;
    p->d    = rho / (unsigned)p->num_dir;                                              /* :52 */
    p->nu   = (float)(1.0 / 12.0 * 1.0 / (p->d * pow((1.0 - p->d), 3.0)) - 1.0 / 8.0); /* :55 */
    p->g    = (float)(SPATIAL_DIM / (SPATIAL_DIM + 2.0) * (1.0 - 2.0 * p->d) / (1.0 - p->d)); /* :58 */
    p->nu_s = p->nu / p->g;                                                            /* :61 */
    p->c_s  = (float)(c / sqrt((double)SPATIAL_DIM));                                  /* :64 */
    p->u    = p->Ma_s * p->c_s;                                                        /* :67 */
}

int lgca_oracle_params_init(lgca_oracle_params* p, int model, const char* tc, float Re, float Ma_s, int cg_radius)
{
    memset(p, 0, sizeof(*p));
    if (lgca_oracle_num_dir(model) < 0) return -1;
    physics(p, model, Re, Ma_s);

    unsigned dim_y;
    if (!strcmp(tc, "pipe")) {
        dim_y = (unsigned)(int)((Re * p->nu_s) / p->u + 0.5); /* :73 */
    } else if (!strcmp(tc, "karman")) {
        float diameter = (Re * p->nu_s) / p->u;               /* :78 */
        dim_y          = (unsigned)(int)(3.0 * diameter + 0.5);
    } else if (!strcmp(tc, "collision")) {
        dim_y = 8;                                            /* :84 */
    } else if (!strcmp(tc, "diffusion") || !strcmp(tc, "periodic") || !strcmp(tc, "box")) {
        dim_y = (unsigned)(int)Re;                            /* :90 */
    } else {
        return -1;                                            /* :94-95 (reference aborts) */
    }
    /* :100 -- always adds between 1 and 2*cg cells */
    dim_y += (2u * (unsigned)cg_radius) - (dim_y % (2u * (unsigned)cg_radius));

    unsigned dim_x;
    if (!strcmp(tc, "pipe") || !strcmp(tc, "karman") || !strcmp(tc, "collision")) dim_x = 2 * dim_y; /* :107 */
    else dim_x = dim_y;                                                                               /* :113 */
    if (!strcmp(tc, "collision")) dim_x++;                                                            /* :122 */

    p->bf_dir = (!strcmp(tc, "pipe") || !strcmp(tc, "karman")) ? 'x' : 0; /* :125-133 */
    p->dim_x  = dim_x;
    p->dim_y  = dim_y;
    finish_dims(p, cg_radius);
    return 0;
}

int lgca_oracle_params_dims(lgca_oracle_params* p, int model, uint32_t dim_x, uint32_t dim_y, int cg_radius,
                            char bf_dir)
{
    memset(p, 0, sizeof(*p));
    if (lgca_oracle_num_dir(model) < 0) return -1;
    physics(p, model, 80.0f, 0.2f);
    p->dim_x  = dim_x;
    p->dim_y  = dim_y;
    p->bf_dir = bf_dir;
    finish_dims(p, cg_radius);
    p->num_cells = (uint64_t)dim_x * dim_y; /* explicit-dims path is not bound to the 32-bit overflow */
    return 0;
}

uint64_t lgca_oracle_initial_forcing(const lgca_oracle_params* p)
{
    return (uint64_t)(0.01 * p->num_cells); /* src/lattice.cpp:450 */
}

uint64_t lgca_oracle_equilibrium_forcing(const lgca_oracle_params* p)
{
    /* src/lattice.cpp:456-460: double expression stored into a float, then ceil(0.5 * N * forcing) */
    float forcing = (float)((8.0 * p->nu_s * p->Ma_s * p->c_s) / pow((double)(float)p->dim_y, 2.0));
    return (uint64_t)ceil(0.5 * p->num_cells * forcing);
}

/* ------------------------------------------------------------------------------------------------
 * Chirality bits, BC painters, initialisers
 * ---------------------------------------------------------------------------------------------- */
void lgca_oracle_fill_rnd(const lgca_oracle_params* p, uint8_t* rnd_bits, lgca_oracle_rng* g)
{
    /* src/lgca_bitset.h:220-224: bit i = rand() % 2, LSB-first inside uint8 blocks (:229-231) */
    memset(rnd_bits, 0, (size_t)((p->num_cells + 7) / 8));
    for (uint64_t i = 0; i < p->num_cells; ++i)
        if (lgca_oracle_rand(g) % 2) rnd_bits[i >> 3] |= (uint8_t)(1u << (i & 7));
}

int lgca_oracle_apply_bc(const lgca_oracle_params* p, const char* name, int32_t* ct)
{
    const uint64_t n  = p->num_cells;
    const uint64_t dx = p->dim_x;
    int edge_type     = -1; /* which solid type goes on which edges */
    int ew            = 0;  /* paint east/west edges too */

    for (uint64_t c = 0; c < n; ++c) ct[c] = LGCA_OR_FLUID; /* apply_cell_type_all, :337-347 */

    if (!strcmp(name, "periodic")) return 0;                /* :221-229 */
    if (!strcmp(name, "pipe") || !strcmp(name, "karman")) edge_type = LGCA_OR_SOLID_NO_SLIP; /* :233-248 */
    else if (!strcmp(name, "reflecting_back"))    { edge_type = LGCA_OR_SOLID_NO_SLIP; ew = 1; } /* :311-334 */
    else if (!strcmp(name, "reflecting_forward")) { edge_type = LGCA_OR_SOLID_SLIP;    ew = 1; }
    else return -1;

    if (ew)
        for (uint64_t c = dx - 1; c < n; c += dx) ct[c] = edge_type; /* east  :350-359 */
    for (uint64_t c = n - dx; c < n; ++c) ct[c] = edge_type;         /* north :362-372 */
    if (ew)
        for (uint64_t c = 0; c < n; c += dx) ct[c] = edge_type;      /* west  :375-384 */
    for (uint64_t c = 0; c < dx; ++c) ct[c] = edge_type;             /* south :387-395 */

    if (!strcmp(name, "karman")) {
        /* :252-280 -- `1 / 10 * m_dim_y` is integer zero; diameter = float(dim_y / 3) with integer
         * division; the distance is rounded to float before the comparison in double. */
        int   center_x = (int)(p->dim_x / 6);
        int   center_y = (int)(p->dim_y / 2 + 1 / 10 * p->dim_y);
        float diameter = (float)(p->dim_y / 3);
        for (uint64_t c = 0; c < n; ++c) {
            int   pos_x = (int)(c % dx);
            int   pos_y = (int)(c / dx);
            float dist  = (float)sqrt(pow((double)(pos_x - center_x), 2.0) + pow((double)(pos_y - center_y), 2.0));
            if (dist < (diameter / 2.0)) ct[c] = LGCA_OR_SOLID_NO_SLIP;
        }
    }
    return 0;
}

static inline int occupy(const lgca_oracle_params* p, lgca_oracle_rng* g)
{
    /* random_uniform() = float(rand()) / float(RAND_MAX), src/utils.h:119-122;
     * compared in double with 1.0 - 1.0 / NUM_DIR, src/lattice.cpp:213 */
    float r = (float)lgca_oracle_rand(g) / (float)2147483647;
    return r > (1.0 - (1.0 / (unsigned)p->num_dir));
}

int lgca_oracle_init(const lgca_oracle_params* p, const char* name, uint8_t* state, const int32_t* ct,
                     lgca_oracle_rng* g)
{
    const uint64_t n = p->num_cells;
    if (!strcmp(name, "zero")) return 0; /* src/lattice.cpp:165-170 */

    if (!strcmp(name, "random")) {       /* src/lattice.cpp:198-217 (serial order = OMP_NUM_THREADS=1) */
        for (uint64_t c = 0; c < n; ++c) {
            if (ct[c] != LGCA_OR_FLUID) continue;
            uint8_t b = state[c];
            for (int d = 0; d < p->num_dir; ++d) {
                if (occupy(p, g)) b |= (uint8_t)(1u << d);
                else              b &= (uint8_t)~(1u << d);
            }
            state[c] = b;
        }
        return 0;
    }
    if (!strcmp(name, "diffusion")) {    /* src/lattice.cpp:465-496 */
        int   center_x = (int)(p->dim_x / 2);
        int   center_y = (int)(p->dim_y / 2);
        float diameter = (float)(p->dim_y / 4);
        for (uint64_t c = 0; c < n; ++c) {
            int   pos_x = (int)(c % p->dim_x);
            int   pos_y = (int)(c / p->dim_x);
            float dist  = (float)sqrt(pow((double)(pos_x - center_x), 2.0) + pow((double)(pos_y - center_y), 2.0));
            if (ct[c] == LGCA_OR_FLUID && dist < (diameter / 2.0)) {
                uint8_t b = state[c];
                for (int d = 0; d < p->num_dir; ++d) {
                    if (occupy(p, g)) b |= (uint8_t)(1u << d);
                    else              b &= (uint8_t)~(1u << d);
                }
                state[c] = b;
            }
        }
        return 0;
    }
    if (!strcmp(name, "single_collision")) { /* src/lattice.cpp:283-308: node index = dir + cell*8 */
        int      inverse_dir = inv_dir(p->model)[0];
        uint64_t n0 = ((uint64_t)(p->dim_x * p->dim_y / 2 + 1)) * 8;
        uint64_t n1 = ((uint64_t)(p->dim_x * p->dim_y / 2 + 5)) * 8 + (uint64_t)inverse_dir;
        state[n0 >> 3] |= (uint8_t)(1u << (n0 & 7));
        state[n1 >> 3] |= (uint8_t)(1u << (n1 & 7));
        return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------------
 * Collision rules
 * ---------------------------------------------------------------------------------------------- */
static void collide_hpp(const uint8_t* in, uint8_t* out)
{
    /* src/lgca_models.h:134-152: arithmetic form on 0/1 values */
    int n0 = in[0], n1 = in[1], n2 = in[2], n3 = in[3];
    int h  = n0 * n2 * (1 - n1) * (1 - n3);
    int v  = n1 * n3 * (1 - n0) * (1 - n2);
    out[0] = (uint8_t)(n0 - h + v);
    out[1] = (uint8_t)(n1 - v + h);
    out[2] = (uint8_t)(n2 - h + v);
    out[3] = (uint8_t)(n3 - v + h);
}

static void collide_fhp(int model, const uint8_t* in, uint8_t* out, int p_in)
{
    /* naming as in the source: a..f = dirs 1..5,0 ; r = rest (src/lgca_models.h:369-374, :570-576) */
    const unsigned a = in[1], b = in[2], c = in[3], d = in[4], e = in[5], f = in[0];
    const unsigned r = (model == LGCA_OR_FHP_I) ? 0u : in[6];
    const unsigned p = p_in ? 1u : 0u, np = p_in ? 0u : 1u; /* (~p)&x with bool p == (!p)&x on 0/1 values */

    /* head-on pairs and symmetric triples: :376-380 / :578-582 / :796-800 */
    unsigned db1    = a & d & ~(b | c | e | f) & 1u;
    unsigned db2    = b & e & ~(a | c | d | f) & 1u;
    unsigned db3    = c & f & ~(a | b | d | e) & 1u;
    unsigned triple = (a ^ b) & (b ^ c) & (c ^ d) & (d ^ e) & (e ^ f);

    unsigned cha = triple | db1 | (p & db2) | (np & db3);
    unsigned chd = cha;
    unsigned chb = triple | db2 | (p & db3) | (np & db1);
    unsigned che = chb;
    unsigned chc = triple | db3 | (p & db1) | (np & db2);
    unsigned chf = chc;
    unsigned chr = 0;

    if (model != LGCA_OR_FHP_I) {
        /* rest particle + one mover -> two movers: :584-589 */
        unsigned ra = r & a & ~(b | c | d | e | f) & 1u;
        unsigned rb = r & b & ~(a | c | d | e | f) & 1u;
        unsigned rc = r & c & ~(a | b | d | e | f) & 1u;
        unsigned rd = r & d & ~(a | b | c | e | f) & 1u;
        unsigned re = r & e & ~(a | b | c | d | f) & 1u;
        unsigned rf = r & f & ~(a | b | c | d | e) & 1u;
        /* two movers 120 degrees apart, no rest -> mover + rest: :591-596 */
        unsigned ra2 = f & b & ~(r | a | c | d | e) & 1u;
        unsigned rb2 = a & c & ~(r | b | d | e | f) & 1u;
        unsigned rc2 = b & d & ~(r | a | c | e | f) & 1u;
        unsigned rd2 = c & e & ~(r | a | b | d | f) & 1u;
        unsigned re2 = d & f & ~(r | a | b | c | e) & 1u;
        unsigned rf2 = e & a & ~(r | b | c | d | f) & 1u;

        cha |= ra | rb | rf | ra2 | rb2 | rf2; /* :598 */
        chd |= rd | rc | re | rd2 | rc2 | re2; /* :599 */
        chb |= rb | ra | rc | rb2 | ra2 | rc2; /* :600 */
        che |= re | rd | rf | re2 | rd2 | rf2; /* :601 */
        chc |= rc | rb | rd | rc2 | rb2 | rd2; /* :602 */
        chf |= rf | ra | re | rf2 | ra2 | re2; /* :603 */
        chr  = ra | rb | rc | rd | re | rf | ra2 | rb2 | rc2 | rd2 | re2 | rf2; /* :604 */

        if (model == LGCA_OR_FHP_III) {
            /* extra FHP-III terms, :816-848.  They are restated in full (and evaluate to zero for
             * every input because db1/db2/db3 already exclude the other four movers). */
            unsigned adbe = db1 & db2 & ~(c | f) & 1u;
            unsigned adcf = db1 & db3 & ~(b | e) & 1u;
            unsigned becf = db2 & db3 & ~(a | d) & 1u;
            unsigned chad = (p & adbe) | (np & adcf) | becf;
            unsigned chbe = (p & becf) | (np & adbe) | adcf;
            unsigned chcf = (p & adcf) | (np & becf) | adbe;
            unsigned adb = db1 & b & ~(c | e | f) & 1u, adc = db1 & c & ~(b | e | f) & 1u;
            unsigned ade = db1 & e & ~(b | c | f) & 1u, adf = db1 & f & ~(b | c | e) & 1u;
            unsigned bea = db2 & a & ~(c | d | f) & 1u, bec = db2 & c & ~(a | d | f) & 1u;
            unsigned bed = db2 & d & ~(a | c | f) & 1u, bef = db2 & f & ~(a | c | d) & 1u;
            unsigned cfa = db3 & a & ~(b | d | e) & 1u, cfb = db3 & b & ~(a | d | e) & 1u;
            unsigned cfd = db3 & d & ~(a | b | e) & 1u, cfe = db3 & e & ~(a | b | d) & 1u;
            unsigned spchad = (adb | ade | adc | adf) | (bec | bef) | (cfb | cfe);
            unsigned spchbe = (bea | bec | bed | bef) | (adc | adf) | (cfa | cfd);
            unsigned spchcf = (cfa | cfb | cfd | cfe) | (adb | ade) | (bea | bed);
            cha |= spchad | chad; chd |= spchad | chad;
            chb |= spchbe | chbe; che |= spchbe | chbe;
            chc |= spchcf | chcf; chf |= spchcf | chcf;
        }
    }
    out[1] = (uint8_t)((a ^ cha) & 1u);
    out[2] = (uint8_t)((b ^ chb) & 1u);
    out[3] = (uint8_t)((c ^ chc) & 1u);
    out[4] = (uint8_t)((d ^ chd) & 1u);
    out[5] = (uint8_t)((e ^ che) & 1u);
    out[0] = (uint8_t)((f ^ chf) & 1u);
    if (model != LGCA_OR_FHP_I) out[6] = (uint8_t)((r ^ chr) & 1u);
}

void lgca_oracle_collide_cell(int model, const uint8_t* in, uint8_t* out, int p)
{
    if (model == LGCA_OR_HPP) collide_hpp(in, out);
    else collide_fhp(model, in, out, p);
}

/* ------------------------------------------------------------------------------------------------
 * Offset tables: src/lgca_models.h:79-132 (HPP), :292-365 (FHP-I), :483-566 (FHP-II), :701-784
 * (FHP-III; identical to FHP-II).  64-bit here -- the reference stores `int`.
 * index: [parity][dir]
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t nb[2][7];    /* offset_to_neighbor_{even,odd} */
    int64_t east[2][7];  /* offset_to_eastern_boundary_*  (applied when the cell is on the WESTERN edge)  */
    int64_t north[2][7]; /* offset_to_northern_boundary_* (applied when the cell is on the SOUTHERN edge) */
    int64_t west[2][7];  /* offset_to_western_boundary_*  (applied when the cell is on the EASTERN edge)  */
    int64_t south[2][7]; /* offset_to_southern_boundary_* (applied when the cell is on the NORTHERN edge) */
} offset_tables;

static void build_tables(int model, int64_t dx, int64_t dy, offset_tables* t)
{
    memset(t, 0, sizeof(*t));
    if (model == LGCA_OR_HPP) {
        for (int par = 0; par < 2; ++par) { /* odd tables are copies of the even ones, :107-131 */
            t->nb[par][0] = 1;  t->nb[par][1] = dx;  t->nb[par][2] = -1;  t->nb[par][3] = -dx;
            t->east[par][2]  = dx;
            t->north[par][3] = dx * dy;
            t->west[par][0]  = -dx;
            t->south[par][1] = -dx * dy;
        }
        return;
    }
    /* even rows, :295-334 */
    t->nb[0][0] = 1; t->nb[0][1] = dx; t->nb[0][2] = dx - 1; t->nb[0][3] = -1; t->nb[0][4] = -dx - 1; t->nb[0][5] = -dx;
    t->east[0][2] = dx; t->east[0][3] = dx; t->east[0][4] = dx;
    t->north[0][4] = dx * dy; t->north[0][5] = dx * dy;
    t->west[0][0] = -dx;
    t->south[0][1] = -dx * dy; t->south[0][2] = -dx * dy + 1;
    /* odd rows, :336-364 */
    t->nb[1][0] = 1; t->nb[1][1] = dx + 1; t->nb[1][2] = dx; t->nb[1][3] = -1; t->nb[1][4] = -dx; t->nb[1][5] = -dx + 1;
    t->east[1][3] = dx;
    t->north[1][4] = dx * dy; t->north[1][5] = dx * dy;
    t->west[1][0] = -dx; t->west[1][1] = -dx; t->west[1][5] = -dx;
    t->south[1][1] = -dx * dy; t->south[1][2] = -dx * dy;
    /* dir 6 (rest particle): all zero, :492,:500,... */
}

/* ------------------------------------------------------------------------------------------------
 * One step: src/omp_lattice.cpp:100-249
 * ---------------------------------------------------------------------------------------------- */
void lgca_oracle_step(const lgca_oracle_params* p, const uint8_t* in, uint8_t* out, const int32_t* ct,
                      const uint8_t* rnd)
{
    const int      nd    = p->num_dir;
    const int64_t  dx    = p->dim_x;
    const int64_t  n     = (int64_t)p->num_cells;
    const int*     INV   = inv_dir(p->model);
    const int*     MIR_X = mir_dir_x(p->model);
    const int*     MIR_Y = mir_dir_y(p->model);
    offset_tables  t;
    build_tables(p->model, dx, (int64_t)p->dim_y, &t);

#pragma omp parallel for schedule(static)
    for (int64_t cell = 0; cell < n; ++cell) {
        const int pos_y = (int)(cell / dx);                   /* :123 */
        const int type  = ct[cell];                           /* :128 */
        const int on_e  = (cell + 1) % dx == 0;               /* :131 */
        const int on_n  = cell >= n - dx;                     /* :132 */
        const int on_w  = cell % dx == 0;                     /* :133 */
        const int on_s  = cell < dx;                          /* :134 */
        const int par   = pos_y % 2 != 0;

        uint8_t node[8] = {0}, tmp[8] = {0};
        for (int dir = 0; dir < nd; ++dir) {                  /* propagation (pull), :141-180 */
            const int inv = INV[dir];
            int64_t   off = t.nb[par][inv];
            if (on_e) off += t.west[par][inv];
            if (on_n) off += t.south[par][inv];
            if (on_w) off += t.east[par][inv];
            if (on_s) off += t.north[par][inv];
            node[dir] = (uint8_t)((in[cell + off] >> dir) & 1u);
        }
        for (int dir = 0; dir < nd; ++dir) tmp[dir] = node[dir]; /* :190 */

        switch (type) {                                       /* :193-231 */
        case LGCA_OR_FLUID:
            lgca_oracle_collide_cell(p->model, node, tmp, (rnd[cell >> 3] >> (cell & 7)) & 1);
            break;
        case LGCA_OR_SOLID_NO_SLIP:
            for (int dir = 0; dir < nd; ++dir) tmp[dir] = node[INV[dir]];
            break;
        case LGCA_OR_SOLID_SLIP:
            if (on_n || on_s)
                for (int dir = 0; dir < nd; ++dir) tmp[dir] = node[MIR_X[dir]];
            if (on_e || on_w)
                for (int dir = 0; dir < nd; ++dir) tmp[dir] = node[MIR_Y[dir]];
            break;
        default: break;
        }
        uint8_t b = 0;                                        /* :235-238 (bits >= NUM_DIR stay 0) */
        for (int dir = 0; dir < nd; ++dir) b |= (uint8_t)((tmp[dir] & 1u) << dir);
        out[cell] = b;
    }
}

void lgca_oracle_steps(const lgca_oracle_params* p, uint8_t* state, uint8_t* scratch, const int32_t* ct,
                       const uint8_t* rnd, int n)
{
    uint8_t* a = state;
    uint8_t* b = scratch;
    for (int i = 0; i < n; ++i) {
        lgca_oracle_step(p, a, b, ct, rnd);
        uint8_t* s = a; a = b; b = s; /* pointer swap, :246-248 */
    }
    if (a != state) memcpy(state, a, (size_t)p->num_cells);
}

/* ------------------------------------------------------------------------------------------------
 * Body force: src/omp_lattice.cpp:254-346
 * ---------------------------------------------------------------------------------------------- */
uint64_t lgca_oracle_body_force(const lgca_oracle_params* p, uint8_t* state, const int32_t* ct, int forcing,
                                lgca_oracle_rng* g, uint32_t* reverted_out)
{
    const uint64_t it_max   = 2 * p->num_cells; /* :258 */
    uint64_t       it       = 0;
    uint32_t       reverted = 0;
    do {
        uint64_t cell = (uint64_t)lgca_oracle_rand(g) % p->num_cells; /* :269 */
        it++;
        if (ct[cell] == LGCA_OR_FLUID) {
            uint8_t s = state[cell], w = s;
            if (p->model == LGCA_OR_HPP) {                               /* :295-310 */
                if (p->bf_dir == 'x' && !(s & 1) && (s & 4))       { w = (uint8_t)((w | 1) & ~4); reverted++; }
                else if (p->bf_dir == 'y' && (s & 2) && !(s & 8))  { w = (uint8_t)((w | 8) & ~2); reverted++; }
            } else {                                                     /* :313-338 */
                if (p->bf_dir == 'x' && !(s & 1) && (s & 8))       { w = (uint8_t)((w | 1) & ~8); reverted++; }
                else if (p->bf_dir == 'y') {
                    if ((s & 2) && !(s & 32)) { w = (uint8_t)((w | 32) & ~2); reverted++; }
                    if ((s & 4) && !(s & 16)) { w = (uint8_t)((w | 16) & ~4); reverted++; }
                }
            }
            state[cell] = w;                                             /* :341 */
        }
    } while (((int64_t)reverted < (int64_t)forcing) && (it < it_max));   /* :345 */
    if (reverted_out) *reverted_out = reverted;
    return it;
}

/* ------------------------------------------------------------------------------------------------
 * Post-processing: src/omp_lattice.cpp:360-394, :397-454, :508-557; src/lattice.cpp:180-195
 * ---------------------------------------------------------------------------------------------- */
void lgca_oracle_cell_post_process(const lgca_oracle_params* p, const uint8_t* s, float* rho, float* mom)
{
    float vx[7], vy[7];
    lattice_vecs(p->model, vx, vy);
    const int64_t n = (int64_t)p->num_cells;
#pragma omp parallel for schedule(static)
    for (int64_t cell = 0; cell < n; ++cell) {
        char  dens = 0;
        float mx = 0.0f, my = 0.0f;
        for (int dir = 0; dir < p->num_dir; ++dir) {
            char ns = (char)((s[cell] >> dir) & 1);
            dens += ns;
            mx += ns * vx[dir];
            my += ns * vy[dir];
        }
        rho[cell]         = (float)dens;
        mom[2 * cell]     = mx;
        mom[2 * cell + 1] = my;
    }
}

void lgca_oracle_mean_post_process(const lgca_oracle_params* p, const float* rho, const float* mom, float* mrho,
                                   float* mmom)
{
    const int      r  = (int)p->cg_radius;
    const uint64_t dx = p->dim_x;
    const int64_t  nc = (int64_t)p->num_coarse_cells;
#pragma omp parallel for schedule(static)
    for (int64_t cc = 0; cc < nc; ++cc) {
        /* anchor = bottom-left cell of the coarse cell, :406-407 */
        const uint64_t cell  = ((uint64_t)cc % p->coarse_dim_x) * (uint64_t)(2 * r)
                             + ((uint64_t)cc / p->coarse_dim_x) * (uint64_t)(2 * r) * dx;
        const int      pos_x = (int)(cell % dx);
        float md = 0.0f, mx = 0.0f, my = 0.0f;
        int   cnt = 0;
        for (int y = 0; y <= 2 * r; ++y) {
            for (int x = 0; x <= 2 * r; ++x) {
                uint64_t nb    = cell + (uint64_t)y * dx + (uint64_t)x;
                int      pos_n = (int)(nb % dx);
                if (nb < p->num_cells && abs(pos_n - pos_x) <= r) { /* :434-436 */
                    cnt++;
                    md += rho[nb];
                    mx += mom[2 * nb];
                    my += mom[2 * nb + 1];
                }
            }
        }
        mrho[cc]         = md / (float)cnt;
        mmom[2 * cc]     = mx / (float)cnt;
        mmom[2 * cc + 1] = my / (float)cnt;
    }
}

void lgca_oracle_mean_velocity(const lgca_oracle_params* p, const int32_t* ct, const float* rho, const float* mom,
                               float out[2])
{
    float    sx = 0.0f, sy = 0.0f;
    uint64_t counter = 0;
    for (uint64_t n = 0; n < p->num_cells; ++n) {
        if (ct[n] == LGCA_OR_FLUID) {
            counter++;
            float dens = rho[n];
            if (dens > 1.0e-06) {
                sx += mom[2 * n] / dens;
                sy += mom[2 * n + 1] / dens;
            }
        }
    }
    out[0] = sx / (float)counter;
    out[1] = sy / (float)counter;
}

uint64_t lgca_oracle_n_particles(const lgca_oracle_params* p, const uint8_t* s)
{
    uint64_t n = 0;
    for (uint64_t c = 0; c < p->num_cells; ++c) n += (uint64_t)__builtin_popcount(s[c]);
    return n;
}

uint64_t lgca_oracle_fnv1a64(const uint8_t* data, uint64_t n)
{
    uint64_t h = 1469598103934665603ull;
    for (uint64_t i = 0; i < n; ++i) { h ^= data[i]; h *= 1099511628211ull; }
    return h;
}
