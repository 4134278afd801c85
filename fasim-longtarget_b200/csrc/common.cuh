// Shared definitions of the B200 triplex-scan engine (device + host side of libfasim_b200.so).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace ltg {

// ---- base codes ------------------------------------------------------------------------------
// DNA base code x: A0 C1 G2 T3, everything else 4 ('N' and any other byte: rules.h:308-311 maps
// every character outside ATGCN to 'N').  Translated DNA / RNA "SSW codes" follow
// ssw_cpp.cpp:13-26: A0 C1 G2 T3 (U -> 0 !) else 4.
constexpr int kBaseOther = 4;

__host__ __device__ inline int dna_code(unsigned char ch)
{
    switch (ch) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return 4;
    }
}
// ssw_cpp.cpp:13-26 (case-insensitive, U -> A)
__host__ __device__ inline int ssw_code(unsigned char ch)
{
    switch (ch) {
    case 'A': case 'a': case 'U': case 'u': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}
// stats.h:306-334 cg_str over nascii (stats.h:201): A1 C2 G3 T4 U5, everything else N(16) -> here 0..5:
// 0 A, 1 C, 2 G, 3 T, 4 U, 5 N
__host__ __device__ inline int stats_code(unsigned char ch)
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 'U': case 'u': return 4;
    default: return 5;
    }
}

// ---- tasks -------------------------------------------------------------------------------------
struct TaskDef {
    int8_t para;      // +1 / -1
    int8_t strand;    // 0 / 1
    int8_t rule;      // 1..18
    int8_t reversed;  // 1: seq2 is the reversed translated segment (ParaMinus, AntiPlus)
    int8_t img[5];    // translated SSW code for base code 0..4 (A,C,G,T,N)
    int8_t comp_src;  // 1: source strand string is complemented (ParaMinus, AntiMinus)
    int8_t pair;      // scan pair that carries this task ...
    int8_t half;      // ... and the 16-bit half of the packed registers it lives in
};
constexpr int kMaxTasks = 48;
constexpr int kMaxPairs = 48;

struct PairDef {
    int16_t task[2];   // indices into the task table (task[1] == task[0] when unpaired)
    int16_t reversed;
    int16_t pad_;
};

// ---- packed 16x2 helpers -----------------------------------------------------------------------
__host__ __device__ inline uint32_t pack16(int lo, int hi) { return (uint32_t)(lo & 0xffff) | ((uint32_t)(hi & 0xffff) << 16); }
__host__ __device__ inline int lo16(uint32_t v) { return (int)(int16_t)(v & 0xffff); }
__host__ __device__ inline int hi16(uint32_t v) { return (int)(int16_t)(v >> 16); }

constexpr int kGapOpen = 16;     // first gap column costs 16 (ssw_cpp.cpp:244, stats.h:947 '\020')
constexpr int kGapExt = 4;       // each further one 4
constexpr int kMatch = 5, kMismatch = -4;
constexpr int kGhost = -16384;   // score of rows beyond the padded query (never reaches a column maximum)
constexpr int kOverflowU8 = 251; // 8-bit scan stops recording once the running maximum reaches 251 (sswNew.cpp:384-396)
constexpr int kQ4CarryF = 132;   // smallest carried F whose decayed value (>= 128) the reference's signed lazy-F test misreads
constexpr int kQ4Guard = 148;    // smallest H that can carry F >= 132 across a stripe boundary

#define LTG_CUDA_CHECK(expr)                                                                       \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ltg::set_error("CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #expr); \
            return LTG_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

void set_error(const char* fmt, ...);

}  // namespace ltg
