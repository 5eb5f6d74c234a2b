// front_util.h — shared by the native file front-ends (png_front.cpp, zip_front.cpp): host threads and result records.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/deft4cu.h"

namespace d4front {

// fn(i) for i in [0, n) on a few host threads (files are independent); D4_HOST_THREADS caps them
template <typename F>
inline void parallel_for(uint32_t n, F&& fn, uint32_t grain = 8) {
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* e = getenv("D4_HOST_THREADS")) hw = (unsigned)atoi(e);
    const unsigned nt = std::max(1u, std::min({hw ? hw : 1u, 32u, (n + grain - 1) / grain}));
    if (nt <= 1) { for (uint32_t i = 0; i < n; i++) fn(i); return; }
    std::atomic<uint32_t> next{0};
    auto body = [&] {
        for (;;) {
            const uint32_t i0 = next.fetch_add(grain);
            if (i0 >= n) break;
            for (uint32_t i = i0; i < std::min(n, i0 + grain); i++) fn(i);
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(body);
    body();
    for (auto& t : th) t.join();
}

// fills n_streams / stream_saved / stream_name / saved_bits of a file result; false when out of memory
inline bool set_streams(deft4cu_file_result& fr, const std::vector<std::string>& names, const std::vector<int64_t>& saved) {
    const size_t ns = names.size();
    fr.n_streams = (uint32_t)ns;
    fr.stream_saved = (int64_t*)calloc(ns ? ns : 1, sizeof(int64_t));
    fr.stream_name = (char**)calloc(ns ? ns : 1, sizeof(char*));
    if (!fr.stream_saved || !fr.stream_name) return false;
    fr.saved_bits = 0;
    for (size_t k = 0; k < ns; k++) {
        fr.stream_saved[k] = saved[k];
        fr.saved_bits += saved[k];
        fr.stream_name[k] = (char*)malloc(names[k].size() + 1);
        if (!fr.stream_name[k]) return false;
        memcpy(fr.stream_name[k], names[k].c_str(), names[k].size() + 1);
    }
    return true;
}

inline void free_result(deft4cu_file_result& r) {
    free(r.out);
    free(r.stream_saved);
    if (r.stream_name) for (uint32_t k = 0; k < r.n_streams; k++) free(r.stream_name[k]);
    free(r.stream_name);
    r.out = nullptr; r.stream_saved = nullptr; r.stream_name = nullptr; r.n_streams = 0;
}

}  // namespace d4front
