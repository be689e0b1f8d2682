// CPU check of vmp::StagePool (voxelmapplus_fastlio2_b200/csrc/vmp_stage.hpp): the helper-thread copy of a scan into the staging
// area and the time-order check of a raw scan that rides along (lio_builder.cpp:75 sorts a scan whose points are out of order).
// Built and run by tests/test_stage_pool.py; prints "ok" or the first failed expectation.
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "../../voxelmapplus_fastlio2_b200/csrc/vmp_stage.hpp"

static int fails = 0;
#define EXPECT(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); fails++; } } while (0)

int main() {
    std::mt19937_64 rng(7);
    for (int threads : {0, 1, 3, 11}) {
        setenv("VMP_COPY_THREADS", std::to_string(threads).c_str(), 1);
        vmp::StagePool pool;
        // plain copies of every size class: below the parallel threshold, odd tails, several MB
        for (size_t bytes : {size_t(0), size_t(1), size_t(4095), size_t(512 * 1024 - 1), size_t(512 * 1024), size_t(600 * 1024 + 7), size_t(3200000), size_t(9999991)}) {
            std::vector<uint8_t> src(bytes + 8), dst(bytes + 8, 0xAB);
            for (auto& b : src) b = (uint8_t)rng();
            EXPECT(pool.copy(dst.data(), src.data(), bytes));
            EXPECT(std::memcmp(dst.data(), src.data(), bytes) == 0);
            for (size_t k = bytes; k < bytes + 8; k++) EXPECT(dst[k] == 0xAB);          // nothing past the end
        }
        // the same copy in two halves (vmp_scan launches the graph between them): the helpers' parts run while the caller is elsewhere
        for (size_t bytes : {size_t(100), size_t(600 * 1024 + 7), size_t(3200000)}) {
            std::vector<uint8_t> src(bytes + 8), dst(bytes + 8, 0xCD);
            for (auto& b : src) b = (uint8_t)rng();
            for (int rep = 0; rep < 20; rep++) {
                pool.begin(dst.data(), src.data(), bytes);
                volatile uint64_t sink = 0;
                for (int k = 0; k < 1000 * (rep % 4); k++) sink += rng();                 // the caller's own work of varying length
                EXPECT(pool.end());
                EXPECT(std::memcmp(dst.data(), src.data(), bytes) == 0);
                src[rep % bytes] ^= 0x5A;
            }
            for (size_t k = bytes; k < bytes + 8; k++) EXPECT(dst[k] == 0xCD);
        }
        // chunked copy: the helpers run through all chunks on their own, the caller picks the chunks up in order
        for (size_t bytes : {size_t(1000), size_t(300 * 1024 + 5), size_t(2400000), size_t(9999991)}) {
            for (size_t chunk : {size_t(384) * 800, size_t(1) << 20, bytes + 1}) {
                std::vector<uint8_t> src(bytes + 8), dst(bytes + 8, 0xEF);
                for (auto& b : src) b = (uint8_t)rng();
                pool.begin_chunks(dst.data(), src.data(), bytes, chunk);
                int c = 0;
                for (size_t off = 0; off < bytes; off += chunk, c++) {
                    pool.wait_chunk(c);
                    const size_t len = std::min(chunk, bytes - off);
                    EXPECT(std::memcmp(dst.data() + off, src.data() + off, len) == 0);         // chunk c is complete once wait_chunk returns
                }
                pool.end_chunks();
                for (size_t k = bytes; k < bytes + 8; k++) EXPECT(dst[k] == 0xEF);
                EXPECT(pool.copy(dst.data(), src.data(), bytes));                              // the other kinds of job still work afterwards
            }
        }
        // x y z out of records of 4 / 8 floats
        for (int stride : {3, 4, 8}) {
            for (int n : {0, 1, 1000, 30000, 200000}) {
                std::vector<float> src((size_t)stride * n + 1), dst(3 * (size_t)n + 1, -7.f);
                for (auto& v : src) v = (float)(rng() % 100000) * 0.01f;
                pool.gather_xyz(dst.data(), src.data(), stride, (size_t)n);
                bool same = true;
                for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) same &= dst[3 * (size_t)i + k] == src[(size_t)stride * i + k];
                EXPECT(same);
                EXPECT(dst[3 * (size_t)n] == -7.f);
            }
        }
        { std::vector<uint8_t> a(700000, 3), b(700000, 0); EXPECT(pool.copy(b.data(), a.data(), a.size())); EXPECT(a == b); }       // a plain copy after a gather
        // raw scans: records of 4 floats, the last one a time offset
        for (int n : {1, 2, 1000, 40000, 200000, 333333}) {
            std::vector<float> src(4 * (size_t)n), dst(4 * (size_t)n);
            for (int i = 0; i < n; i++) { src[4 * i] = (float)i; src[4 * i + 1] = 1.f; src[4 * i + 2] = 2.f; src[4 * i + 3] = (float)(i / 3) * 0.01f; }   // non-decreasing, with ties
            EXPECT(pool.copy(dst.data(), src.data(), src.size() * 4, 4));
            EXPECT(dst == src);
            if (n >= 2) {
                // one inversion anywhere (also across the boundary of two helpers' parts) is found
                for (int trial = 0; trial < 6; trial++) {
                    const int k = trial == 0 ? 1 : trial == 1 ? n - 1 : 1 + (int)(rng() % (uint64_t)(n - 1));
                    const float keep = src[4 * k + 3];
                    src[4 * k + 3] = src[4 * (k - 1) + 3] - 1.0f;
                    EXPECT(!pool.copy(dst.data(), src.data(), src.size() * 4, 4));
                    EXPECT(dst == src);                                             // the copy itself is complete either way
                    src[4 * k + 3] = keep;
                }
                EXPECT(pool.copy(dst.data(), src.data(), src.size() * 4, 4));
            }
        }
    }
    if (!fails) std::printf("ok\n");
    return fails ? 1 : 0;
}
