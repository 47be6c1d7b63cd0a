// Bit-sliced collision / wall rules of the four lattice-gas models, one machine word = 32 sites
// (multi-spin coding).  Every function works on whole words; bit j of every word belongs to the same
// lattice site.  Written once for device (LOP3 via inline PTX) and host (plain C++, used only by the
// CPU unit test that checks these networks exhaustively against the oracle's truth tables).
//
// Reference semantics (paths relative to /root/reference):
//   HPP      ModelDescriptor<HPP>::collide      src/lgca_models.h:134-152
//   FHP-I    ModelDescriptor<FHP_I>::collide    src/lgca_models.h:367-395
//   FHP-II   ModelDescriptor<FHP_II>::collide   src/lgca_models.h:568-613
//   FHP-III  ModelDescriptor<FHP_III>::collide  src/lgca_models.h:786-856  (its extra terms :816-848
//            are identically zero, so the table equals FHP-II's -- SURVEY.md fact 7 / A.3)
//   walls    bounce_back / bounce_forward_x / bounce_forward_y, e.g. src/lgca_models.h:397-427
//
// The networks are NOT transcriptions of the reference's formulas: they are re-derived from the
// collision table to minimise 3-input logic ops (the integer pipe is the co-limiter of this kernel):
// the movers of the two triangles (even / odd directions) are counted with one full adder each, which
// classifies every site (one mover, head-on candidate, 120-degree pair, alternating triple) in a few ops;
// the rest-particle rules of FHP-II collapse to "flip a trio of adjacent directions and the rest bit".
// 31 LOP3 per 32 sites for FHP-II/III (the reference's formulas lifted verbatim compile to 76).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LGCA_HD __host__ __device__ __forceinline__
#else
#define LGCA_HD inline
#endif

namespace lgca_b200 {

enum : int { MODEL_HPP = 0, MODEL_FHP_I = 1, MODEL_FHP_II = 2, MODEL_FHP_III = 3 };

LGCA_HD constexpr int num_dir_of(int model) { return model == MODEL_HPP ? 4 : (model == MODEL_FHP_I ? 6 : 7); }
// FHP-III as coded in the reference has FHP-II's table: both run the same network.
LGCA_HD constexpr int rule_of(int model) { return model == MODEL_FHP_III ? MODEL_FHP_II : model; }

// LOP3 truth-table operands: build a LUT as an expression of these, e.g. (TA & TB) | TC
constexpr uint32_t TA = 0xF0, TB = 0xCC, TC = 0xAA;

template <uint32_t LUT>
LGCA_HD uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT & 0xFF));
    return r;
#else
    uint32_t r = 0;
    for (int m = 0; m < 8; ++m) {
        if (!((LUT >> m) & 1u)) continue;
        uint32_t ta = (m & 4) ? a : ~a;
        uint32_t tb = (m & 2) ? b : ~b;
        uint32_t tc = (m & 1) ? c : ~c;
        r |= ta & tb & tc;
    }
    return r;
#endif
}

constexpr uint32_t LUT_XOR3 = TA ^ TB ^ TC;
constexpr uint32_t LUT_MAJ  = (TA & TB) | (TA & TC) | (TB & TC);
constexpr uint32_t LUT_OR3  = TA | TB | TC;
constexpr uint32_t LUT_AND3 = TA & TB & TC;
constexpr uint32_t LUT_MUX  = (TA & TB) | (~TA & TC);   // a ? b : c
constexpr uint32_t LUT_XOR_OR = TA ^ (TB | TC);         // a ^ (b | c)

// ---------------------------------------------------------------------------------------------
// HPP: head-on pairs rotate by 90 degrees.  Changed states: 0101 <-> 1010, i.e. the four bits
// alternate: (n0^n1)&(n1^n2)&(n2^n3).
// ---------------------------------------------------------------------------------------------
LGCA_HD void collide_hpp(uint32_t (&n)[7])
{
    uint32_t t = lop3<(TA ^ TB) & (TB ^ TC)>(n[0], n[1], n[2]);
    uint32_t c = lop3<TA & (TB ^ TC)>(t, n[2], n[3]);
    n[0] ^= c; n[1] ^= c; n[2] ^= c; n[3] ^= c;
}

// ---------------------------------------------------------------------------------------------
// FHP family.  Directions: 0=E 1=NE 2=NW 3=W 4=SW 5=SE (6=rest); i and i+3 are opposite.
//   p = chirality word (frozen per-site random bit, src/omp_lattice.cpp:84,198):
//   head-on pair (i,i+3), all other movers empty  -> rotates to (i+1,i+4) if p==0, (i-1,i+2) if p==1
//   symmetric triple (0,2,4)<->(1,3,5)
//   FHP-II: rest + single mover c             -> movers c-1,c+1 (rest consumed)
//           movers c-1,c+1 only, no rest      -> mover c + rest
//
// Network.  The six movers are counted per TRIANGLE -- even directions (0,2,4) and odd directions (1,3,5) --
// with one full adder each:
//     s_e,c_e = XOR3/MAJ(n0,n2,n4)     s_o,c_o = XOR3/MAJ(n1,n3,n5)
// which classifies every site with a handful of 3-input ops:
//     head-on pair somewhere       HO  = s_e & s_o & ~c_e & (n0 == n3) & (n1 == n4)
//     alternating triple           tri = (s_e == c_e) & (s_e^s_o) & (c_e^c_o)
// Pair (j, j+3) flips with T_j = tri | (HO & ~(p ? n_{j+2} : n_{j+1})): under HO exactly one pair is full and
// only the pair it does NOT rotate onto stays unchanged (no per-pair head-on signals are needed).
// Rest rules (FHP-II/III).  Both fire only when ONE triangle is empty: "rest + single mover c" (the other
// triangle holds exactly one mover) and "movers c-1, c+1, no rest" (the other triangle holds exactly two).
// In both, the occupied triangle is cleared of what it holds, and every vertex v of the EMPTY triangle is set
// unless the vertex opposite to it (v+3, in the occupied triangle) is occupied.  Hence, with
//     EE = even triangle empty & (r ? odd has one : odd has two),   EO = likewise with the roles swapped,
// direction i of the even triangle becomes   EE ? ~n_{i+3} : (~EO & (n_i ^ T_i))   and symmetrically for odd i;
// the rest bit flips when EE | EO.  31 LOP3 in total for FHP-II/III, 20 for FHP-I (counted in the SASS).
// ---------------------------------------------------------------------------------------------
template <bool WITH_REST>
LGCA_HD void collide_fhp(uint32_t (&n)[7], uint32_t p)
{
    const uint32_t se = lop3<LUT_XOR3>(n[0], n[2], n[4]);
    const uint32_t ce = lop3<LUT_MAJ>(n[0], n[2], n[4]);
    const uint32_t so = lop3<LUT_XOR3>(n[1], n[3], n[5]);
    const uint32_t co = lop3<LUT_MAJ>(n[1], n[3], n[5]);
    // head-on pair somewhere: one even and one odd mover (s_e & s_o, no carry) sitting opposite each other.
    // With s_e & s_o, "pairs (0,3) and (1,4) are each empty or full" leaves exactly the three head-on states and
    // the all-six state, which ~c_e removes.  3 ops, no per-pair head-on signals.
    const uint32_t u0 = lop3<TA & ~(TB ^ TC) & 0xFF>(se, n[0], n[3]);
    const uint32_t u1 = lop3<TA & ~(TB ^ TC) & 0xFF>(so, n[1], n[4]);
    const uint32_t HO = lop3<TA & TB & ~TC & 0xFF>(u0, u1, ce);
    // symmetric triple: one triangle full (s = c = 1), the other empty (s = c = 0)
    //   tri = (s_e == c_e) & (s_e ^ s_o) & (c_e ^ c_o)      -- 4 inputs, 2 ops
    const uint32_t t1  = lop3<~(TA ^ TB) & (TA ^ TC) & 0xFF>(se, ce, so);
    const uint32_t tri = lop3<TA & (TB ^ TC)>(t1, ce, co);
    // per-pair change masks.  Under HO exactly one pair k is full (n_k = n_{k+3} = 1, everything else empty) and it
    // rotates onto pair k+1 (p = 0) or k-1 (p = 1): pair j flips unless it is the third pair, i.e. unless
    // pair j+1 (p = 0) / pair j+2 (p = 1) is the full one:   T_j = tri | (HO & ~(p ? n_{j+2} : n_{j+1}))
    // (reference: db1 = dirs 1,4; db2 = dirs 2,5; db3 = dirs 3,0).  Written out that is 6 ops; the 5-op form below
    // was found by exhaustive search over LOP3 networks (scripts/lop3_search.c, block "T"):
    //   nc  = no collision;  q = tri ? ~n0 : (p ^ n0);  T_1, T_2 from (n4 | n5, nc, q);  T_0 = f(tri, T_1, T_2)
    const uint32_t nc  = lop3<0x03>(tri, HO, p);        // ~tri & ~HO
    const uint32_t q   = lop3<0x56>(tri, p, n[0]);
    const uint32_t t14 = lop3<0x32>(n[4], nc, q);
    const uint32_t t25 = lop3<0x31>(n[5], nc, q);
    const uint32_t t30 = lop3<0x86>(tri, t14, t25);

    if (WITH_REST) {
        const uint32_t r  = n[6];
        // fo: the odd triangle holds exactly one mover (r set) or exactly two (r clear); fe likewise for the even one
        constexpr uint32_t LUT_F = (TA & TB & ~TC) | (~TA & ~TB & TC & 0xFF); // a ? (b & ~c) : (~b & c)
        const uint32_t fo = lop3<LUT_F>(r, so, co);
        const uint32_t fe = lop3<LUT_F>(r, se, ce);
        const uint32_t EE = lop3<TA & ~TB & ~TC & 0xFF>(fo, se, ce); // even triangle empty, odd one triggers a rest rule
        const uint32_t EO = lop3<TA & ~TB & ~TC & 0xFF>(fe, so, co);
        uint32_t o[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const uint32_t T    = (i == 1 || i == 4) ? t14 : ((i == 2 || i == 5) ? t25 : t30);
            const uint32_t own  = (i & 1) ? EO : EE;   // this direction's triangle is the empty one
            const uint32_t oth  = (i & 1) ? EE : EO;   // ... or the one being cleared
            const uint32_t w    = lop3<~TA & (TB ^ TC) & 0xFF>(oth, n[i], T);         // ~oth & (n_i ^ T)
            o[i] = lop3<(TA & ~TB) | (~TA & TC)>(own, n[(i + 3) % 6], w);             // own ? ~n_{i+3} : w
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) n[i] = o[i];
        n[6] = lop3<TA ^ (TB | TC)>(r, EE, EO);
    } else {
        n[1] ^= t14; n[4] ^= t14;
        n[2] ^= t25; n[5] ^= t25;
        n[3] ^= t30; n[0] ^= t30;
    }
}

template <int MODEL>
LGCA_HD void collide(uint32_t (&n)[7], uint32_t p)
{
    if (rule_of(MODEL) == MODEL_HPP) collide_hpp(n);
    else if (rule_of(MODEL) == MODEL_FHP_I) collide_fhp<false>(n, p);
    else collide_fhp<true>(n, p);
}

// ---------------------------------------------------------------------------------------------
// Direction permutations of the wall rules (src/lgca_models.h:44-48 HPP, :239-243 / :446-450 FHP)
// ---------------------------------------------------------------------------------------------
template <int MODEL> LGCA_HD constexpr int inv_dir(int d)
{
    return rule_of(MODEL) == MODEL_HPP ? ((d + 2) & 3) : (d == 6 ? 6 : (d + 3) % 6);
}
template <int MODEL> LGCA_HD constexpr int mir_x_dir(int d) // mirror at the x axis (N/S walls)
{
    return rule_of(MODEL) == MODEL_HPP ? ((4 - d) & 3) : (d == 6 ? 6 : (6 - d) % 6);
}
template <int MODEL> LGCA_HD constexpr int mir_y_dir(int d) // mirror at the y axis (E/W walls)
{
    return rule_of(MODEL) == MODEL_HPP ? ((d & 1) ? d : (d ^ 2)) : (d == 6 ? 6 : (9 - d) % 6);
}

// Wall handling of one word of sites (SURVEY A.4 / src/omp_lattice.cpp:193-231), applied on top of the
// collided values n[] given the streamed-in (pre-collision) values in[]:
//   no-slip    -> out[d] = in[INV d]
//   slip       -> MIR_Y on E/W edge columns, else MIR_X on N/S edge rows, else pass-through
// `ns`/`sl` are the solid masks, `ew` the mask of sites on the E/W domain edge, `ns_row` all-ones when
// the row is the northern or southern domain edge.
template <int MODEL, bool HAS_NS, bool HAS_SL>
LGCA_HD void apply_walls(uint32_t (&n)[7], const uint32_t (&in)[7], uint32_t ns, uint32_t sl, uint32_t ew, uint32_t ns_row)
{
    constexpr int ND = num_dir_of(MODEL);
    if (HAS_NS) {
#pragma unroll
        for (int d = 0; d < ND; ++d) n[d] = lop3<LUT_MUX>(ns, in[inv_dir<MODEL>(d)], n[d]);
    }
    if (HAS_SL) {
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            uint32_t s = lop3<LUT_MUX>(ns_row, in[mir_x_dir<MODEL>(d)], in[d]);
            s          = lop3<LUT_MUX>(ew, in[mir_y_dir<MODEL>(d)], s);
            n[d]       = lop3<LUT_MUX>(sl, s, n[d]);
        }
    }
}

// Collision + walls: fluid sites collide, solid sites reflect.
template <int MODEL, bool HAS_NS, bool HAS_SL>
LGCA_HD void collide_and_walls(uint32_t (&n)[7], uint32_t p, uint32_t ns, uint32_t sl, uint32_t ew, uint32_t ns_row)
{
    constexpr int ND = num_dir_of(MODEL);
    uint32_t in[7];
#pragma unroll
    for (int d = 0; d < 7; ++d) in[d] = d < ND ? n[d] : 0u;
    collide<MODEL>(n, p);
    apply_walls<MODEL, HAS_NS, HAS_SL>(n, in, ns, sl, ew, ns_row);
}

} // namespace lgca_b200
