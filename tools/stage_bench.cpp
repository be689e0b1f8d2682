// Host-side micro-benchmark of the scan staging copy (vmp_stage.hpp), 2.4 MB = one C2 scan: mode 0 four pool.copy calls (one barrier per chunk),
// mode 1 begin_chunks / wait_chunk (helpers run through all chunks), mode 2 one pool.copy.  g++ -O2 -std=c++17 -pthread tools/stage_bench.cpp
#include <chrono>
#include <cstdio>
#include <vector>
#include "../voxelmapplus_fastlio2_b200/csrc/vmp_stage.hpp"
int main() {
    vmp::StagePool pool;
    const size_t bytes = 2400000;
    std::vector<char> src(bytes, 1), dst(bytes);
    auto now = [] { return std::chrono::steady_clock::now(); };
    for (int mode = 0; mode < 3; mode++) {
        double best = 1e9, sum = 0;
        for (int rep = 0; rep < 200; rep++) {
            for (auto& c : src) c++;          // dirty the source like a fresh scan would be (in cache though)
            auto t0 = now();
            if (mode == 0) { const size_t per = 600000; for (size_t off = 0; off < bytes; off += per) pool.copy(dst.data() + off, src.data() + off, std::min(per, bytes - off)); }
            else if (mode == 1) { const size_t per = 300288; pool.begin_chunks(dst.data(), src.data(), bytes, per); int c = 0; for (size_t off = 0; off < bytes; off += per, c++) pool.wait_chunk(c); pool.end_chunks(); }
            else pool.copy(dst.data(), src.data(), bytes);
            double us = std::chrono::duration<double, std::micro>(now() - t0).count();
            if (rep >= 20) { best = std::min(best, us); sum += us; }
        }
        std::printf("mode %d: best %.1f us, mean %.1f us\n", mode, best, sum / 180);
    }
}
