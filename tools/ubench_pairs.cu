// Pairwise pipe-sharing microbenchmark for sm_100a: interleaves two instruction kinds 1:1 on
// independent register chains and reports the combined issue rate.  If two kinds share an
// execution pipe the combined rate is the harmonic combination; if not, they overlap.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_pairs tools/ubench_pairs.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ILP = 4;          // chains per kind
constexpr int ITERS = 2048;
constexpr int UNROLL = 16;

enum Kind { K_X = 0, K_XR, K_M3, K_VA, K_M2, K_IMAD, K_IADD, K_LOP, K_PRMT, K_SHFL, K_LDS, K_NONE, K_COUNT };
static const char* kname[K_COUNT] = { "viaddmnmx", "viaddmnmx_relu", "vimnmx3", "viadd16x2", "vimnmx2", "imad", "iadd3",
                                      "lop3", "prmt", "shfl", "lds32", "none" };

template <int K>
__device__ __forceinline__ uint32_t apply(uint32_t a, uint32_t b, uint32_t c, const uint32_t* sm)
{
    if (K == K_X)    return __viaddmax_s16x2(a, b, c);
    if (K == K_XR)   return __viaddmax_s16x2_relu(a, b, c);
    if (K == K_M3)   return __vimax3_s16x2(a, b, c);
    if (K == K_VA)   return __vadd2(a, b);
    if (K == K_M2)   { uint32_t r; asm volatile("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
    if (K == K_IMAD) return a * b + c;
    if (K == K_IADD) { uint32_t r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
    if (K == K_LOP)  { uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
    if (K == K_PRMT) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c & 0x7777u)); return r; }
    if (K == K_SHFL) return __shfl_up_sync(0xffffffffu, a, 1);
    if (K == K_LDS)  return sm[a & 1023];
    return a;
}

template <int KA, int KB>
__global__ void __launch_bounds__(256) k_pair(uint32_t* out, uint32_t seed, int iters, unsigned long long* clk)
{
    __shared__ uint32_t sm[1024];
    uint32_t a[ILP], b[ILP], c[ILP], d[ILP];
    const uint32_t t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed * (i + 1) + t; d[i] = seed * (i + 7) ^ t;
        b[i] = (seed >> 3) + i * 0x00010001u + (t & 3);
        c[i] = seed ^ (t * 2654435761u + i);
    }
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i * 7 + seed;
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                a[i] = apply<KA>(a[i], b[i], c[i], sm);
                if (KB != K_NONE) d[i] = apply<KB>(d[i], b[i], c[i], sm);
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= a[i] ^ d[i];
    out[t] = r;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static uint32_t* d_out; static unsigned long long* d_clk; static int nsm;

template <int KA, int KB>
static int run(bool first)
{
    const int blocks = nsm * 4, threads = 256;
    k_pair<KA, KB><<<blocks, threads>>>(d_out, 12345u, 32, d_clk);
    CK(cudaDeviceSynchronize());
    unsigned long long best = ~0ull;
    static unsigned long long hclk[4096];
    for (int rep = 0; rep < 3; ++rep) {
        k_pair<KA, KB><<<blocks, threads>>>(d_out, 777u + rep, ITERS, d_clk);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hclk, d_clk, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (int i = 0; i < blocks; ++i) if (hclk[i] > mx) mx = hclk[i];
        if (mx < best) best = mx;
    }
    double per_kind = (double)ITERS * UNROLL * ILP * blocks * threads;
    double n_kinds = (KB == K_NONE) ? 1.0 : 2.0;
    printf("%s\"%s+%s\": %.2f", first ? "" : ", ", kname[KA], kname[KB], per_kind * n_kinds / (double)best / nsm);
    return 0;
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    nsm = p.multiProcessorCount;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * nsm * 4 * 256));
    CK(cudaMalloc(&d_clk, sizeof(unsigned long long) * 4096));
    printf("{\"unit\": \"thread-instr/clk/SM, both kinds summed\", \"pairs\": {");
#define R(A, B, F) if (run<A, B>(F)) return 1;
    R(K_X, K_NONE, true) R(K_M3, K_NONE, false) R(K_VA, K_NONE, false) R(K_M2, K_NONE, false) R(K_IMAD, K_NONE, false)
    R(K_IADD, K_NONE, false) R(K_LOP, K_NONE, false) R(K_PRMT, K_NONE, false) R(K_SHFL, K_NONE, false) R(K_LDS, K_NONE, false)
    R(K_X, K_XR, false) R(K_X, K_M3, false) R(K_X, K_VA, false) R(K_X, K_M2, false) R(K_X, K_IMAD, false)
    R(K_X, K_IADD, false) R(K_X, K_LOP, false) R(K_X, K_PRMT, false) R(K_X, K_SHFL, false) R(K_X, K_LDS, false)
    R(K_VA, K_IMAD, false) R(K_VA, K_M2, false) R(K_VA, K_M3, false) R(K_VA, K_IADD, false) R(K_VA, K_LOP, false)
    R(K_M2, K_IMAD, false) R(K_M2, K_M3, false) R(K_M2, K_IADD, false) R(K_M2, K_LOP, false) R(K_M2, K_M2, false)
    R(K_M3, K_IMAD, false) R(K_IADD, K_IMAD, false) R(K_LOP, K_IMAD, false) R(K_IADD, K_LOP, false) R(K_PRMT, K_IMAD, false)
    printf("}}\n");
    return 0;
}
