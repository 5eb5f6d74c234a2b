// checksum.cuh — CRC-32 and Adler-32 of every stream's decoded bytes on the device (SURVEY.md §8f row 1:
// GZFile.java:130-145 and ZLibFile.java:42-51 recompute them from getUncompressedData(); doing it here keeps
// the 2.6x-larger decoded data off PCIe).
//
// Both checksums are linear in the data once the init/xor-out is set aside, so they are computed in pieces:
//   CRC  : raw(A ‖ B) = raw(A) * x^(8|B|) + raw(B) over GF(2)[x]/P.  The decoded pool is cut on a 1 KiB address
//          grid; a CTA takes CK_NT consecutive cells of one stream, every thread folds one cell with a
//          slice-by-4 table in shared memory (aligned 128-bit loads), the cells are combined by a butterfly
//          with the constant multipliers x^(8*1024*2^k), the result is advanced to the end of the stream
//          (x^(8*after), a product of precomputed x^(8*2^k) taken by one warp) and XOR-ed into the stream's
//          accumulator.  A stream's head (before its first cell boundary) only lacks leading bytes, which a
//          raw CRC does not see; its tail (after the last full cell) is folded serially by the last CTA.
//   Adler: A = 1 + sum(bytes), B = n + sum((n - i) * byte_i): per-cell sums plus (cell sum) * (bytes after the
//          cell), accumulated with 64-bit atomics and reduced mod 65521 at the end.
#pragma once
#include "common.cuh"

namespace d4 {

constexpr int CK_NT = 128;           // threads per CTA = cells per job
constexpr int CK_CELL_LOG2 = 10;     // 1 KiB cells
constexpr uint32_t CK_CELL = 1u << CK_CELL_LOG2;

__device__ uint32_t g_xpow8[48];     // x^(8 * 2^k) mod P, reflected (filled by k_crc_init)

__device__ __forceinline__ uint32_t crc_mulmod(uint32_t a, uint32_t b) {  // GF(2) polynomial product mod P (reflected)
    uint32_t r = 0;
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        if (a & 0x80000000u) r ^= b;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1) ? 0xEDB88320u : 0);
    }
    return r;
}
__global__ void k_crc_init() {
    uint32_t p = 0x00800000u;  // x^8
    for (int k = 0; k < 48; k++) { g_xpow8[k] = p; p = crc_mulmod(p, p); }
}
// x^(8n) mod P by one warp: lane k contributes x^(8*2^k) when bit k of n is set
__device__ __forceinline__ uint32_t crc_xpow8n_warp(uint64_t n, int lane) {
    uint32_t v = ((n >> lane) & 1) ? g_xpow8[lane] : 0x80000000u;
    if (lane < 16 && ((n >> (32 + lane)) & 1)) v = crc_mulmod(v, g_xpow8[32 + lane]);
    for (int d = 16; d > 0; d >>= 1) v = crc_mulmod(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

struct CkJob { uint32_t stream, chunk; };
struct CkAcc { unsigned long long a, b; uint32_t crc; uint32_t pad; };

__global__ void __launch_bounds__(CK_NT)
k_checksum_cells(const uint8_t* __restrict__ out, const uint64_t* __restrict__ off, const uint64_t* __restrict__ len,
                 const CkJob* __restrict__ jobs, CkAcc* __restrict__ acc) {
    __shared__ uint32_t T[4][256];
    __shared__ uint32_t s_w[CK_NT / 32];
    __shared__ unsigned long long s_a[CK_NT / 32], s_b[CK_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < 256; i += CK_NT) {
        uint32_t c = (uint32_t)i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? 0xEDB88320u : 0);
        T[0][i] = c;
    }
    __syncthreads();
    for (int k = 1; k < 4; k++) {
        for (int i = tid; i < 256; i += CK_NT) { uint32_t p = T[k - 1][i]; T[k][i] = (p >> 8) ^ T[0][p & 0xff]; }
        __syncthreads();
    }
    const CkJob job = jobs[blockIdx.x];
    const uint64_t B = off[job.stream], E = B + len[job.stream];
    const uint64_t cellFirst = B >> CK_CELL_LOG2, cellEnd = E >> CK_CELL_LOG2;  // cells [cellFirst, cellEnd) end inside the stream
    const uint64_t c0 = cellFirst + (uint64_t)job.chunk * CK_NT;
    uint64_t c1 = c0 + CK_NT;
    if (c1 > cellEnd) c1 = cellEnd;
    const int m = c1 > c0 ? (int)(c1 - c0) : 0;      // cells of this job
    const bool last = c1 == cellEnd || m == 0;       // this job also folds the stream's tail
    // thread t <-> cell c0 + t - (CK_NT - m): the job's last cell sits in the last thread, so every cell is
    // followed by whole cells only and the butterfly multipliers are constants
    uint32_t crc = 0;
    unsigned long long ta = 0, tb = 0;
    const int ci = tid - (CK_NT - m);
    if (ci >= 0) {
        const uint64_t cell = c0 + (uint64_t)ci;
        uint64_t lo = cell << CK_CELL_LOG2;
        const uint64_t hi = lo + CK_CELL;
        uint32_t a = 0, b = 0;
        if (lo < B) {  // the stream's head: bytewise
            for (uint64_t k = B; k < hi; k++) {
                const uint32_t v = out[k];
                crc = T[0][(crc ^ v) & 0xff] ^ (crc >> 8);
                a += v; b += a;
            }
        } else {
            const uint4* p = (const uint4*)(out + lo);
#pragma unroll 2
            for (int k = 0; k < (int)(CK_CELL / 16); k++) {
                const uint4 q = __ldg(p + k);
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t x = w[j];
                    crc ^= x;
                    crc = T[3][crc & 0xff] ^ T[2][(crc >> 8) & 0xff] ^ T[1][(crc >> 16) & 0xff] ^ T[0][crc >> 24];
                    const uint32_t b0 = x & 0xff, b1 = (x >> 8) & 0xff, b2 = (x >> 16) & 0xff, b3 = x >> 24;
                    b += 4 * a + 4 * b0 + 3 * b1 + 2 * b2 + b3;
                    a += b0 + b1 + b2 + b3;
                }
            }
        }
        const uint64_t after = E - hi;
        ta = a;
        tb = ((unsigned long long)(b % 65521u) + (unsigned long long)(a % 65521u) * (after % 65521u)) % 65521u;
    }
    // butterfly: after level k every group of 2^(k+1) threads holds the raw CRC of its cells
    uint32_t v = crc;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t mine = ((lane >> k) & 1) ? v : crc_mulmod(v, g_xpow8[CK_CELL_LOG2 + k]);
        v = mine ^ __shfl_xor_sync(0xffffffffu, mine, 1 << k);
    }
    for (int d = 16; d > 0; d >>= 1) { ta += __shfl_xor_sync(0xffffffffu, ta, d); tb += __shfl_xor_sync(0xffffffffu, tb, d); }
    if (lane == 0) { s_w[wid] = v; s_a[wid] = ta; s_b[wid] = tb; }
    __syncthreads();
    if (wid == 0) {
        uint32_t r = lane < CK_NT / 32 ? s_w[lane] : 0;
        unsigned long long A = lane < CK_NT / 32 ? s_a[lane] : 0, Bs = lane < CK_NT / 32 ? s_b[lane] : 0;
#pragma unroll
        for (int k = 0; (1 << k) < CK_NT / 32; k++) {
            const uint32_t mine = ((lane >> k) & 1) ? r : crc_mulmod(r, g_xpow8[CK_CELL_LOG2 + 5 + k]);
            r = mine ^ __shfl_xor_sync(0xffffffffu, mine, 1 << k);
        }
        for (int d = 16; d > 0; d >>= 1) { A += __shfl_xor_sync(0xffffffffu, A, d); Bs += __shfl_xor_sync(0xffffffffu, Bs, d); }
        uint64_t done = m ? (c1 << CK_CELL_LOG2) : B;   // the CRC in r covers the stream up to here
        if (last) {  // tail: the bytes after the last whole cell (or the whole of a stream inside one cell)
            if (lane == 0) {
                uint32_t a = 0, b = 0, c = r;
                uint64_t k = done < B ? B : done;
                while (k < E) {
                    uint64_t stop = k + 4096 < E ? k + 4096 : E;
                    for (; k < stop; k++) {
                        const uint32_t x = out[k];
                        c = T[0][(c ^ x) & 0xff] ^ (c >> 8);
                        a += x; b += a;
                    }
                    a %= 65521u; b %= 65521u;
                }
                r = c; A += a; Bs += b;
            }
            r = __shfl_sync(0xffffffffu, r, 0);
            done = E;
        }
        const uint32_t adv = crc_xpow8n_warp(E - done, lane);
        if (lane == 0) {
            atomicXor(&acc[job.stream].crc, crc_mulmod(r, adv));
            atomicAdd(&acc[job.stream].a, A);
            atomicAdd(&acc[job.stream].b, Bs % 65521u);
        }
    }
}

// one warp per stream: init / xor-out of the CRC, the constant terms of Adler-32
__global__ void k_checksum_final(const uint64_t* __restrict__ len, const CkAcc* __restrict__ acc, uint32_t n,
                                 uint32_t* __restrict__ crc_out, uint32_t* __restrict__ adler_out) {
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n) return;
    const uint64_t L = len[s];
    const uint32_t adv = crc_xpow8n_warp(L, lane);
    if (lane == 0) {
        crc_out[s] = ~(acc[s].crc ^ crc_mulmod(0xFFFFFFFFu, adv));
        const uint32_t A = (uint32_t)((1 + acc[s].a) % 65521u);
        const uint32_t Bv = (uint32_t)((L % 65521u + acc[s].b) % 65521u);
        adler_out[s] = (Bv << 16) | A;
    }
}

}  // namespace d4
