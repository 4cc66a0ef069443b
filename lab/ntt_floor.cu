// Minimum-operation-count bound for the coset LDE launch ("prove it with the minimum-op-count lab kernel").
// The kernels below issue EXACTLY the field arithmetic an n = 2^14 coset transform needs — one pre-scale multiply per element,
// 14 butterfly levels (every twiddle inside a 16- or 64-point group a shift: field.cuh f_shl), and the fewest table-twiddle
// multiplies a radix-16 (3 stages) or radix-64 (2 stages) decomposition allows — on registers only: no shared memory, no
// barrier, no index arithmetic beyond one coalesced load and store per element. They do NOT exchange data between threads, so
// the output is not a transform; the instruction stream and its dependencies are those of a transform whose data movement is
// free. Their throughput is therefore an upper bound for any kernel built from this arithmetic; the product kernel
// (lde_block_kernel_t<3, 1024, 14>) is measured beside them on the same shape (135 columns x 8 cosets x 2^14).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I zk-circuits_b200/csrc -o lab/ntt_floor lab/ntt_floor.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "field.cuh"
namespace zkb { __device__ u64 d_rootA[2048], d_rootB[2048], d_rootC[1024]; }   // tables ntt.cuh expects from kernels.cu
#include "ntt.cuh"
using namespace zkb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// MODE 0: copy only (the traffic bound). MODE 1: radix-16 count (1 + 3 x 15/16 multiplies). MODE 2: radix-64 count
// (1 + 2 multiplies, two extra shift-twiddle stages). MODE 3: butterflies only, no multiplies at all.
template <int MODE>
__global__ void __launch_bounds__(256) floor_kernel(const u64* __restrict__ coeffs, const u64* __restrict__ prescale, u64* __restrict__ out,
                                                    unsigned n) {
    const unsigned col = blockIdx.y, jb = blockIdx.z;
    const u64* src = coeffs + (size_t)col * n;
    const u64* ps = prescale + (size_t)jb * n;
    u64* dst = out + ((size_t)col * 8 + jb) * n;
    const unsigned i0 = blockIdx.x * (256 * 16) + threadIdx.x;
    u64 r[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) r[e] = src[i0 + e * 256];
    if (MODE != 0) {
        if (MODE != 3) {
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] = f_mul(r[e], __ldg(ps + i0 + e * 256));
        }
        const unsigned tmask = (1u << NTT_SM_LG) - 1;
        if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                u64 tw[16];
                if (MODE == 1) {
#pragma unroll
                    for (int p = 1; p < 16; ++p) tw[p] = __ldg(&d_W14[((i0 + pass) * p * 64) & tmask]);
                }
                RadixStep<4, false, 0, 0, 0>::run(r);
                if (MODE == 1) {
#pragma unroll
                    for (int p = 1; p < 16; ++p) r[p] = f_mul(r[p], tw[p]);
                }
            }
            RadixStep<2, false, 0, 0, 0>::run(r); RadixStep<2, false, 0, 0, 0>::run(r + 4);
            RadixStep<2, false, 0, 0, 0>::run(r + 8); RadixStep<2, false, 0, 0, 0>::run(r + 12);
        } else {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                u64 tw[16];
#pragma unroll
                for (int p = 0; p < 16; ++p) tw[p] = __ldg(&d_W14[((i0 + pass) * (p + 1) * 64) & tmask]);
                RadixStep<4, false, 0, 0, 0>::run(r);                   // levels 1-4 of the 64-point group
                // the 64-point group's middle twiddles w_64^(a q): compile-time shifts (a = this 16-point slice's residue)
                r[1] = f_mul_w64<3>(r[1]); r[2] = f_mul_w64<6>(r[2]); r[3] = f_mul_w64<9>(r[3]); r[5] = f_mul_w64<2>(r[5]);
                r[6] = f_mul_w64<4>(r[6]); r[7] = f_mul_w64<6>(r[7]); r[9] = f_mul_w64<1>(r[9]); r[10] = f_mul_w64<2>(r[10]);
                r[11] = f_mul_w64<3>(r[11]); r[13] = f_mul_w64<21>(r[13]); r[14] = f_mul_w64<42>(r[14]); r[15] = f_mul_w64<63>(r[15]);
                RadixStep<2, false, 0, 0, 0>::run(r); RadixStep<2, false, 0, 0, 0>::run(r + 4);      // levels 5-6
                RadixStep<2, false, 0, 0, 0>::run(r + 8); RadixStep<2, false, 0, 0, 0>::run(r + 12);
#pragma unroll
                for (int p = 0; p < 16; ++p) r[p] = f_mul(r[p], tw[p]);                                // 63/64 of the elements
            }
            RadixStep<2, false, 0, 0, 0>::run(r); RadixStep<2, false, 0, 0, 0>::run(r + 4);
            RadixStep<2, false, 0, 0, 0>::run(r + 8); RadixStep<2, false, 0, 0, 0>::run(r + 12);
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) dst[i0 + e * 256] = r[e];
}

template <int MODE>
static float time_mode(const u64* c, const u64* ps, u64* out, unsigned n, int cols, int reps) {
    dim3 grid(n / (256 * 16), cols, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) floor_kernel<MODE><<<grid, 256>>>(c, ps, out, n);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) floor_kernel<MODE><<<grid, 256>>>(c, ps, out, n);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const unsigned n = 1u << 14;
    const int cols = 135, reps = 20;
    u64 *c, *ps, *out;
    CK(cudaMalloc(&c, sizeof(u64) * n * cols));
    CK(cudaMalloc(&ps, sizeof(u64) * n * 8));
    CK(cudaMalloc(&out, sizeof(u64) * n * cols * 8));
    std::vector<u64> h((size_t)n * cols);
    u64 x = 88172645463325252ull;
    for (auto& v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x % GL_P; }
    CK(cudaMemcpy(c, h.data(), sizeof(u64) * n * cols, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ps, h.data(), sizeof(u64) * n * 8, cudaMemcpyHostToDevice));
    std::vector<u64> w(n);
    for (auto& v : w) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x % GL_P; }
    CK(cudaMemcpyToSymbol(d_W14, w.data(), sizeof(u64) * n));
    const double bytes = 72.0 * n * cols;         // read n, write 8 n words per column: the LDE launch's algorithmic traffic
    const char* names[4] = {"copy only (traffic bound)", "radix-16 operation count (1 + 2.8 multiplies, 14 levels)",
                            "radix-64 operation count (1 + 2 multiplies, 2 shift stages, 14 levels)", "butterflies only (14 levels, no multiply)"};
    float ms[4] = {time_mode<0>(c, ps, out, n, cols, reps), time_mode<1>(c, ps, out, n, cols, reps),
                   time_mode<2>(c, ps, out, n, cols, reps), time_mode<3>(c, ps, out, n, cols, reps)};
    CK(cudaDeviceSynchronize());
    for (int m = 0; m < 4; ++m) printf("%-76s %.4f ms  %7.1f GB/s algorithmic\n", names[m], ms[m], bytes / (ms[m] * 1e-3) / 1e9);
    return 0;
}
