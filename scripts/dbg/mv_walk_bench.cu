// Build (no GPU needed): nvcc -O3 -std=c++17 -Xcompiler -O3 -o /tmp/mv_bench scripts/dbg/mv_walk_bench.cu -Llgca_b200 -llgca_b200 -Xlinker -rpath=$PWD/lgca_b200
// Run: /tmp/mv_bench <class bytes file> <dim_x> <rows> <model 0..3>   (class bytes: state byte of fluid cells, 0 elsewhere)
// CPU timing harness of the ordered mean-velocity walk (host code of lgca_mv.cu), no GPU needed.
#include "../../lgca_b200/csrc/lgca_mv.cu"
#include <chrono>
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv)
{
    const char* path = argv[1]; const uint32_t dx = atoi(argv[2]), rows = atoi(argv[3]); const int model = atoi(argv[4]);
    std::vector<uint8_t> cls((size_t)dx * rows);
    FILE* f = fopen(path, "rb"); if (!f || fread(cls.data(), 1, cls.size(), f) != cls.size()) { puts("read failed"); return 1; } fclose(f);
    const MvTables& T = mv_tables(model);
    const uint32_t spr = (dx + MV_SEG_CELLS - 1) / MV_SEG_CELLS; const size_t nseg = (size_t)spr * rows;
    std::vector<int32_t> rec(nseg * MV_REC);
    for (uint32_t row = 0; row < rows; ++row) for (uint32_t k = 0; k < spr; ++k) {
        const uint32_t x0 = k * MV_SEG_CELLS, n = std::min<uint32_t>(MV_SEG_CELLS, dx - x0);
        mv_summarise_host(T, cls.data() + (size_t)row * dx + x0, n, rec.data(), (size_t)row * spr + k, nseg);
    }
    double best = 1e9; float sums[2]; uint64_t st[2];
    for (int it = 0; it < 15; ++it) {
        sums[0] = sums[1] = 0; st[0] = st[1] = 0;
        auto t0 = std::chrono::steady_clock::now();
        mv_walk(T, rec.data(), cls.data(), dx, rows, sums, st);
        best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    printf("%ux%u model %d: walk %.3f ms  fast %llu walked %llu  sums %.9g %.9g (bits %08x %08x)\n", dx, rows, model, best,
           (unsigned long long)st[0], (unsigned long long)st[1], sums[0], sums[1], *(uint32_t*)&sums[0], *(uint32_t*)&sums[1]);
    return 0;
}
