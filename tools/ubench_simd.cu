// Integer-SIMD issue-rate microbenchmark for sm_100a.
//
// Measures the sustained per-SM issue rate (thread-instructions / clk / SM) of the
// packed 16-bit DPX-style instructions the scan kernel is built from
// (VIADDMNMX.S16x2[.RELU], VIMNMX.S16x2, VIMNMX3.S16x2, VIADD.16x2) plus IMAD / IADD3 /
// LOP3 / SHFL / LDS for pipe-sharing experiments.  Register-only, ILP 8, 32 warps/SM.
// Output: one JSON object on stdout; bench.py reads profiles/int_simd_peak.json made from it.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_simd tools/ubench_simd.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ILP = 8;
constexpr int ITERS = 4096;
constexpr int UNROLL = 16;

enum Op { OP_VIADDMNMX, OP_VIADDMNMX_RELU, OP_VIMNMX, OP_VIMNMX3, OP_VIADD, OP_IADD3, OP_IMAD, OP_LOP3,
          OP_MIX_DP6, OP_MIX_DP6_IMAD2, OP_MIX_DP55, OP_SHFL, OP_LDS128, OP_MIX_DP55_LDS, OP_COUNT };

static const char* op_name[OP_COUNT] = {
    "viaddmnmx_s16x2", "viaddmnmx_s16x2_relu", "vimnmx_s16x2", "vimnmx3_s16x2", "viadd_16x2", "iadd3", "imad", "lop3",
    "mix_dp6", "mix_dp6_plus_2imad", "mix_dp5.5", "shfl", "lds128", "mix_dp5.5_plus_lds" };
// instructions of the class under test per inner body (per chain element)
static const double op_instr[OP_COUNT] = { 1, 1, 1, 1, 1, 1, 1, 1, 6, 6, 5.5, 1, 1, 5.5 };

template <int OP>
__global__ void __launch_bounds__(256) k_bench(uint32_t* out, uint32_t seed, int iters, unsigned long long* clk)
{
    __shared__ uint4 sm[256 * 2];
    uint32_t a[ILP], b[ILP], c[ILP];
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed * (i + 1) + t;
        b[i] = (seed >> 3) + i * 0x00010001u;
        c[i] = seed ^ (t * 2654435761u + i);
    }
    sm[threadIdx.x] = make_uint4(a[0], b[0], c[0], a[1]);
    sm[threadIdx.x + 256] = make_uint4(a[2], b[2], c[2], a[3]);
    __syncthreads();
    const uint32_t mge = 0xFFFCFFFCu;   // (-4,-4)
    const uint32_t mgo = 0xFFF0FFF0u;   // (-16,-16)
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == OP_VIADDMNMX)            a[i] = __viaddmax_s16x2(a[i], b[i], c[i]);
                else if (OP == OP_VIADDMNMX_RELU)  a[i] = __viaddmax_s16x2_relu(a[i], b[i], c[i]);
                else if (OP == OP_VIMNMX)          a[i] = __vmaxs2(a[i], b[i]);
                else if (OP == OP_VIMNMX3)         a[i] = __vimax3_s16x2(a[i], b[i], c[i]);
                else if (OP == OP_VIADD)           a[i] = __vadd2(a[i], b[i]);
                else if (OP == OP_IADD3)           a[i] = a[i] + b[i] + c[i];
                else if (OP == OP_IMAD)            a[i] = a[i] * b[i] + c[i];
                else if (OP == OP_LOP3)            a[i] = (a[i] & b[i]) ^ c[i];
                else if (OP == OP_SHFL)            a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
                else if (OP == OP_LDS128) {
                    uint4 v = sm[(a[i] & 255) + ((u & 1) << 8)];
                    a[i] = v.x ^ v.w;   // 1 LDS.128 + 1 LOP3 per element
                }
                else if (OP == OP_MIX_DP6 || OP == OP_MIX_DP6_IMAD2) {
                    // a=Hd/H, b=E, c=F ; per cell-pair: 6 packed ops (reference formulation)
                    uint32_t s = 0x00050005u;
                    uint32_t tt = __viaddmax_s16x2_relu(a[i], s, b[i]);
                    uint32_t h = __vmaxs2(tt, c[i]);
                    uint32_t cm = __vmaxs2(a[(i + 1) % ILP], h);
                    uint32_t uu = __vadd2(h, mgo);
                    b[i] = __viaddmax_s16x2(b[i], mge, uu);
                    c[i] = __viaddmax_s16x2(c[i], mge, uu);
                    a[i] = cm;
                    if (OP == OP_MIX_DP6_IMAD2) {   // two extra FMA-pipe ops riding along
                        b[i] = b[i] * 3u + seed;
                        c[i] = c[i] * 5u + seed;
                    }
                }
                else if (OP == OP_MIX_DP55 || OP == OP_MIX_DP55_LDS) {
                    // t-formulation, 2 rows per body (11 packed ops per 2 cell-pairs)
                    if ((i & 1) == 0) {
                        uint32_t s0 = 0x00050005u, s1 = 0xFFFCFFFCu;
                        if (OP == OP_MIX_DP55_LDS) {
                            uint4 v = sm[(threadIdx.x) + ((u & 1) << 8)];
                            s0 = v.x; s1 = v.y;
                        }
                        uint32_t t0_ = __viaddmax_s16x2_relu(a[i], s0, b[i]);
                        uint32_t t1_ = __viaddmax_s16x2_relu(a[i + 1], s1, b[i + 1]);
                        uint32_t u0 = __vadd2(t0_, mgo);
                        uint32_t u1 = __vadd2(t1_, mgo);
                        b[i] = __viaddmax_s16x2(b[i], mge, u0);
                        b[i + 1] = __viaddmax_s16x2(b[i + 1], mge, u1);
                        uint32_t f0 = c[i];
                        uint32_t f1 = __viaddmax_s16x2(f0, mge, u0);
                        c[i] = __viaddmax_s16x2(f1, mge, u1);
                        a[i] = __vmaxs2(t0_, f0);
                        a[i + 1] = __vmaxs2(t1_, f1);
                        c[i + 1] = __vimax3_s16x2(c[i + 1], t0_, t1_);
                    }
                }
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= a[i] ^ b[i] ^ c[i];
    out[t] = r;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
static int run(int nsm, uint32_t* d_out, unsigned long long* d_clk, double* rate_out, double* clk_mhz_out, double* ms_out)
{
    const int blocks = nsm * 4, threads = 256;      // 32 warps / SM
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_bench<OP><<<blocks, threads>>>(d_out, 12345u, 64, d_clk);          // warm-up
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    unsigned long long best_clk = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        k_bench<OP><<<blocks, threads>>>(d_out, 12345u + rep, ITERS, d_clk);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        unsigned long long hclk[2048];
        CK(cudaMemcpy(hclk, d_clk, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (int i = 0; i < blocks; ++i) if (hclk[i] > mx) mx = hclk[i];
        if (ms < best_ms) { best_ms = ms; best_clk = mx; }
    }
    double div = (OP == OP_MIX_DP55 || OP == OP_MIX_DP55_LDS) ? 2.0 : 1.0;   // body runs on half the chain slots
    double thread_instr = (double)ITERS * UNROLL * ILP / div * op_instr[OP] * (OP == OP_MIX_DP55 || OP == OP_MIX_DP55_LDS ? 2.0 : 1.0);
    // for DP55: per (i even) body = 11 ops = 2 * 5.5; ILP/2 bodies per u  -> ITERS*UNROLL*(ILP/2)*11
    double total = thread_instr * (double)blocks * threads;
    *rate_out = total / (double)best_clk / nsm;          // thread-instr / clk / SM (clk = slowest block's cycles)
    *clk_mhz_out = (double)best_clk / (best_ms * 1e3);
    *ms_out = best_ms;
    return 0;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    uint32_t* d_out; unsigned long long* d_clk;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * nsm * 4 * 256));
    CK(cudaMalloc(&d_clk, sizeof(unsigned long long) * 2048));
    double rate[OP_COUNT], mhz[OP_COUNT], ms[OP_COUNT];
#define RUN(OPX) if (run<OPX>(nsm, d_out, d_clk, &rate[OPX], &mhz[OPX], &ms[OPX])) return 1;
    RUN(OP_VIADDMNMX) RUN(OP_VIADDMNMX_RELU) RUN(OP_VIMNMX) RUN(OP_VIMNMX3) RUN(OP_VIADD) RUN(OP_IADD3) RUN(OP_IMAD)
    RUN(OP_LOP3) RUN(OP_MIX_DP6) RUN(OP_MIX_DP6_IMAD2) RUN(OP_MIX_DP55) RUN(OP_SHFL) RUN(OP_LDS128) RUN(OP_MIX_DP55_LDS)
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_attr\": %d, \"ops\": {", p.name, nsm, p.clockRate);
    for (int i = 0; i < OP_COUNT; ++i)
        printf("%s\"%s\": {\"thread_instr_per_clk_per_sm\": %.2f, \"ms\": %.3f, \"sm_mhz_observed\": %.0f}",
               i ? ", " : "", op_name[i], rate[i], ms[i], mhz[i]);
    printf("}}\n");
    return 0;
}
