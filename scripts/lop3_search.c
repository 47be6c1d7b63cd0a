// Development helper: exhaustive search for LOP3 (3-input LUT) networks of sub-blocks of the FHP-II collision
// function (lgca_b200/csrc/lgca_collide.cuh).  Every signal is a 256-bit truth table over the eight primary inputs
// (n0..n5, rest, chirality), so don't-cares of a block (input combinations that cannot occur) are handled for free.
//
//   gcc -O2 -o /tmp/lop3_search scripts/lop3_search.c && /tmp/lop3_search <block> <max_gates>
//
// A block names its input signals and target signals; the search asks whether all targets can be produced with
// <= max_gates gates (targets may be produced in complemented form: consumers are LOP3s, complement is free).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t w[4]; } tt;

static tt tt_not(tt a) { for (int i = 0; i < 4; ++i) a.w[i] = ~a.w[i]; return a; }
static tt tt_and(tt a, tt b) { for (int i = 0; i < 4; ++i) a.w[i] &= b.w[i]; return a; }
static tt tt_or(tt a, tt b) { for (int i = 0; i < 4; ++i) a.w[i] |= b.w[i]; return a; }
static tt tt_xor(tt a, tt b) { for (int i = 0; i < 4; ++i) a.w[i] ^= b.w[i]; return a; }
static int tt_zero(tt a) { return !(a.w[0] | a.w[1] | a.w[2] | a.w[3]); }
static int tt_eq(tt a, tt b) { return tt_zero(tt_xor(a, b)); }
static tt tt_const(int v) { tt a; for (int i = 0; i < 4; ++i) a.w[i] = v ? ~0ull : 0ull; return a; }
static tt tt_mux(tt s, tt a, tt b) { return tt_or(tt_and(s, a), tt_and(tt_not(s), b)); }

static tt prim(int bit) // primary input `bit` of the 8-bit index
{
    tt a = tt_const(0);
    for (int i = 0; i < 256; ++i)
        if ((i >> bit) & 1) a.w[i >> 6] |= 1ull << (i & 63);
    return a;
}

// is t a function of (x, y, z)?  (also true when t is constant or depends on fewer)
static int func_of(tt t, tt x, tt y, tt z)
{
    tt nt = tt_not(t);
    for (int c = 0; c < 8; ++c) {
        tt m = tt_and(tt_and((c & 4) ? x : tt_not(x), (c & 2) ? y : tt_not(y)), (c & 1) ? z : tt_not(z));
        if (!tt_zero(tt_and(m, t)) && !tt_zero(tt_and(m, nt))) return 0;
    }
    return 1;
}

static tt lut3(int lut, tt a, tt b, tt c)
{
    tt r = tt_const(0);
    for (int m = 0; m < 8; ++m)
        if ((lut >> m) & 1)
            r = tt_or(r, tt_and(tt_and((m & 4) ? a : tt_not(a), (m & 2) ? b : tt_not(b)), (m & 1) ? c : tt_not(c)));
    return r;
}

#define MAXS 40
static tt   sig[MAXS];
static char name[MAXS][48];
static int  nsig;
static tt   target[8];
static char tname[8][16];
static int  ntarget;
static long visited;

static int known(tt t)
{
    for (int i = 0; i < nsig; ++i)
        if (tt_eq(sig[i], t) || tt_eq(sig[i], tt_not(t))) return 1;
    return 0;
}

static int search(int gates_left, unsigned done_mask)
{
    // realise every target that is now a single gate away (never hurts: it needs its own gate anyway)
    int base_nsig = nsig;
    int progress = 1;
    while (progress) {
        progress = 0;
        for (int t = 0; t < ntarget; ++t) {
            if (done_mask & (1u << t)) continue;
            int ok = 0, bi = 0, bj = 0, bk = 0;
            if (known(target[t])) { done_mask |= 1u << t; progress = 1; continue; }
            for (int i = 0; i < nsig && !ok; ++i)
                for (int j = i; j < nsig && !ok; ++j)
                    for (int k = j; k < nsig && !ok; ++k)
                        if (func_of(target[t], sig[i], sig[j], sig[k])) { ok = 1; bi = i; bj = j; bk = k; }
            if (ok) {
                if (gates_left == 0) { nsig = base_nsig; return 0; }
                --gates_left;
                sig[nsig] = target[t];
                {   // LUT of the realised target over its triple (unreachable input classes filled with 0)
                    int lut = 0;
                    for (int c = 0; c < 8; ++c) {
                        tt m = tt_and(tt_and((c & 4) ? sig[bi] : tt_not(sig[bi]), (c & 2) ? sig[bj] : tt_not(sig[bj])),
                                      (c & 1) ? sig[bk] : tt_not(sig[bk]));
                        if (!tt_zero(tt_and(m, target[t]))) lut |= 1 << c;
                    }
                    snprintf(name[nsig], sizeof name[nsig], "%s=lut%02x(%d,%d,%d)", tname[t], lut, bi, bj, bk);
                }
                ++nsig;
                done_mask |= 1u << t;
                progress = 1;
            }
        }
    }
    int remaining = 0;
    for (int t = 0; t < ntarget; ++t) remaining += !(done_mask & (1u << t));
    if (remaining == 0) {
        printf("FOUND with signals:\n");
        for (int i = 0; i < nsig; ++i) printf("  [%d] %s\n", i, name[i]);
        fflush(stdout);
        nsig = base_nsig;
        return 1;
    }
    if (gates_left <= remaining) { nsig = base_nsig; return 0; } // every remaining target needs its own gate + >= 1 helper
    // add one helper gate: any LUT of any triple, deduplicated by function
    int n = nsig;
    static tt seen_stack[8][60000];
    tt* seen = seen_stack[gates_left];
    int nseen = 0;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            for (int k = j + 1; k < n; ++k)
                for (int lut = 1; lut < 128; ++lut) { // complement classes: lut and ~lut give complementary signals
                    tt g = lut3(lut, sig[i], sig[j], sig[k]);
                    if (known(g)) continue;
                    int dup = 0;
                    for (int s = 0; s < nseen && !dup; ++s) dup = tt_eq(seen[s], g) || tt_eq(seen[s], tt_not(g));
                    if (dup) continue;
                    if (nseen < 60000) seen[nseen++] = g;
                    sig[nsig] = g;
                    snprintf(name[nsig], sizeof name[nsig], "g=lut%02x(%d,%d,%d)", lut, i, j, k);
                    ++nsig;
                    ++visited;
                    int r = search(gates_left - 1, done_mask);
                    --nsig;
                    if (r) { nsig = base_nsig; return 1; }
                }
    nsig = base_nsig;
    return 0;
}

static void add_sig(const char* nm, tt t) { sig[nsig] = t; snprintf(name[nsig], sizeof name[nsig], "%s", nm); ++nsig; }
static void add_target(const char* nm, tt t) { target[ntarget] = t; snprintf(tname[ntarget], sizeof tname[ntarget], "%s", nm); ++ntarget; }

int main(int argc, char** argv)
{
    const char* block = argc > 1 ? argv[1] : "class";
    int max_gates = argc > 2 ? atoi(argv[2]) : 5;
    tt n[6], r = prim(6), p = prim(7);
    for (int i = 0; i < 6; ++i) n[i] = prim(i);
    // the signals of the shipped network
    tt se = tt_xor(tt_xor(n[0], n[2]), n[4]), so = tt_xor(tt_xor(n[1], n[3]), n[5]);
    tt ce = tt_or(tt_or(tt_and(n[0], n[2]), tt_and(n[0], n[4])), tt_and(n[2], n[4]));
    tt co = tt_or(tt_or(tt_and(n[1], n[3]), tt_and(n[1], n[5])), tt_and(n[3], n[5]));
    tt ze = tt_and(tt_not(se), tt_not(ce)), zo = tt_and(tt_not(so), tt_not(co));
    tt tri = tt_or(tt_and(tt_and(se, ce), zo), tt_and(tt_and(so, co), ze));
    tt HO = tt_and(tt_and(tt_and(se, so), tt_not(ce)),
                   tt_and(tt_not(tt_xor(n[0], n[3])), tt_not(tt_xor(n[1], n[4]))));
    tt fo = tt_mux(r, tt_and(so, tt_not(co)), tt_and(tt_not(so), co));
    tt fe = tt_mux(r, tt_and(se, tt_not(ce)), tt_and(tt_not(se), ce));
    tt EE = tt_and(ze, fo), EO = tt_and(zo, fe);
    tt T[3]; // T[j]: pair (j, j+3) flips
    for (int j = 0; j < 3; ++j) T[j] = tt_or(tri, tt_and(HO, tt_not(tt_mux(p, n[(j + 2) % 3], n[(j + 1) % 3]))));
    tt out[7];
    for (int i = 0; i < 6; ++i) {
        tt own = (i & 1) ? EO : EE, oth = (i & 1) ? EE : EO;
        out[i] = tt_mux(own, tt_not(n[(i + 3) % 6]), tt_and(tt_not(oth), tt_xor(n[i], T[i % 3])));
    }
    out[6] = tt_xor(r, tt_or(EE, EO));

    if (!strcmp(block, "class")) { // (se,ce,so,co,r) -> tri, EE, EO   [shipped: 6 gates]
        add_sig("se", se); add_sig("ce", ce); add_sig("so", so); add_sig("co", co); add_sig("r", r);
        add_target("tri", tri); add_target("EE", EE); add_target("EO", EO);
    } else if (!strcmp(block, "restsig")) { // (se,ce,so,co,r) -> EE, EO, rest output   [shipped: 5 gates]
        add_sig("se", se); add_sig("ce", ce); add_sig("so", so); add_sig("co", co); add_sig("r", r);
        add_target("EE", EE); add_target("EO", EO); add_target("o6", out[6]);
    } else if (!strcmp(block, "class2")) { // + HO and the rest output   [shipped: 6 + 3 + 1]
        add_sig("se", se); add_sig("ce", ce); add_sig("so", so); add_sig("co", co); add_sig("r", r);
        add_sig("n0", n[0]); add_sig("n3", n[3]); add_sig("n1", n[1]); add_sig("n4", n[4]);
        add_target("tri", tri); add_target("EE", EE); add_target("EO", EO); add_target("HO", HO);
    } else if (!strcmp(block, "T")) { // (tri, HO, p, n0, n1, n2) -> T0, T1, T2   [shipped: 6 gates]
        add_sig("tri", tri); add_sig("HO", HO); add_sig("p", p);
        for (int i = 0; i < 6; ++i) { char b[8]; snprintf(b, 8, "n%d", i); add_sig(b, n[i]); }
        add_target("T0", T[0]); add_target("T1", T[1]); add_target("T2", T[2]);
    } else if (!strcmp(block, "pair")) { // (a, b, T, EE, EO) -> a', b'   [shipped: 4 gates]
        add_sig("a", n[0]); add_sig("b", n[3]); add_sig("T0", T[0]); add_sig("EE", EE); add_sig("EO", EO);
        add_target("o0", out[0]); add_target("o3", out[3]);
    } else if (!strcmp(block, "pairT")) { // pair outputs straight from tri/HO/p   [shipped: 2 + 4 gates]
        add_sig("a", n[0]); add_sig("b", n[3]); add_sig("tri", tri); add_sig("HO", HO); add_sig("p", p);
        add_sig("n1", n[1]); add_sig("n2", n[2]); add_sig("EE", EE); add_sig("EO", EO);
        add_target("o0", out[0]); add_target("o3", out[3]);
    } else if (!strcmp(block, "pair1") || !strcmp(block, "pair2")) { // pair outputs with the T gate folded in [shipped: 1 + 4]
        const int j = block[4] - '0';
        tt nc = tt_and(tt_not(tri), tt_not(HO));
        tt q  = tt_mux(tri, tt_not(n[0]), tt_xor(p, n[0]));
        char b[8];
        snprintf(b, 8, "n%d", j); add_sig(b, n[j]);
        snprintf(b, 8, "n%d", j + 3); add_sig(b, n[j + 3]);
        add_sig("nc", nc); add_sig("q", q); add_sig("EE", EE); add_sig("EO", EO);
        snprintf(b, 8, "o%d", j); add_target(b, out[j]);
        snprintf(b, 8, "o%d", j + 3); add_target(b, out[j + 3]);
    } else if (!strcmp(block, "pair0")) { // pair (0,3) from tri, T1, T2 [shipped: 1 + 4]
        add_sig("n0", n[0]); add_sig("n3", n[3]); add_sig("tri", tri); add_sig("T1", T[1]); add_sig("T2", T[2]);
        add_sig("EE", EE); add_sig("EO", EO);
        add_target("o0", out[0]); add_target("o3", out[3]);
    } else if (!strcmp(block, "rest")) { // everything after the T's: 6 movers + rest   [shipped: 4 + 12 + 1 = 17]
        for (int i = 0; i < 6; ++i) { char b[8]; snprintf(b, 8, "n%d", i); add_sig(b, n[i]); }
        add_sig("r", r); add_sig("se", se); add_sig("ce", ce); add_sig("so", so); add_sig("co", co);
        add_sig("T0", T[0]); add_sig("T1", T[1]); add_sig("T2", T[2]);
        for (int i = 0; i < 7; ++i) { char b[8]; snprintf(b, 8, "o%d", i); add_target(b, out[i]); }
    } else {
        fprintf(stderr, "unknown block\n");
        return 2;
    }
    printf("block %s: %d inputs, %d targets, <= %d gates\n", block, nsig, ntarget, max_gates);
    int ok = search(max_gates, 0);
    printf("%s (%ld helper gates tried)\n", ok ? "feasible" : "NOT feasible", visited);
    return ok ? 0 : 1;
}
