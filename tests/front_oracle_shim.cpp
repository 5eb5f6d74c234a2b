// Test infrastructure: deft4cu_optimise_batch implemented over the CPU oracle, so that the host-side file front-ends
// (deft4j_b200/csrc/png_front.cpp) can be checked against the reference's golden files without a GPU.  Linked only
// into tests/_build/libfront_oracle.so by tests/hosttest_lib.py; never part of libdeft4cu.so.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/deft4cu.h"
#include "../oracle/deft_oracle.h"

static uint32_t adler32(const uint8_t* p, size_t n) {
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < n; i++) { a = (a + p[i]) % 65521; b = (b + a) % 65521; }
    return (b << 16) | a;
}

static uint32_t crc32_of(const uint8_t* p, size_t n) {
    uint32_t c = ~0u;
    for (size_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1)));
    }
    return ~c;
}

extern "C" {
int deft4cu_optimise_batch(const uint8_t* const* in, const uint64_t* in_len, uint32_t n, uint32_t flags, deft4cu_result* results) {
    for (uint32_t i = 0; i < n; i++) {
        deft4cu_result& r = results[i];
        memset(&r, 0, sizeof r);
        size_t consumed = 0;
        ora_stream* s = ora_parse(in[i], in_len[i], &consumed);
        if (!s) { r.status = DEFT4CU_ERR_PARSE; continue; }
        r.consumed_bytes = consumed;
        r.size_bits_in = ora_size_bits(s);
        r.saved_bits = ora_optimise(s, (flags & DEFT4CU_MERGE_BLOCKS) ? 1 : 0);
        r.size_bits_out = ora_size_bits(s);
        const size_t cap = in_len[i] + 64;
        r.out = (uint8_t*)malloc(cap);
        r.out_len = ora_write(s, r.out, cap);
        const size_t u = ora_uncompressed_len(s);
        std::vector<uint8_t> data(u ? u : 1);
        ora_uncompressed(s, data.data());
        r.uncompressed_len = u;
        r.adler32 = adler32(data.data(), u);
        r.crc32 = crc32_of(data.data(), u);
        ora_free(s);
    }
    return DEFT4CU_OK;
}
void deft4cu_free_results(deft4cu_result* results, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) { free(results[i].out); results[i].out = nullptr; }
}
}
