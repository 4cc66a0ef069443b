// sm_100a kernels for the Plonky2 proving hot path: Poseidon/Merkle (a3, a4),
// coset LDE-NTT (a2, a5), partial products (a7), quotient (a8), openings (a9), FRI (a10-a13).
// Integer-only; no tensor cores (nothing here is a dense contraction). Launchers are in kernels.h.
#include "kernels.h"
#include "poseidon.cuh"
#include <map>
#include <tuple>
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace zkb {

#define ZKB_CUDA_CHECK(x)                                                                                   \
    do {                                                                                                    \
        cudaError_t e_ = (x);                                                                               \
        if (e_ != cudaSuccess)                                                                              \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #x);     \
    } while (0)

static std::atomic<unsigned long long> g_kernel_launches{0};
unsigned long long kernel_launch_count() { return g_kernel_launches.load(); }
void kernel_launch_count_add(unsigned long long n) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }
#define ZKB_COUNT_LAUNCH() (g_kernel_launches.fetch_add(1, std::memory_order_relaxed))

// ---------------------------------------------------------------------------------------------
// device tables
// ---------------------------------------------------------------------------------------------
// T = 2^32-th root of unity; T^E = rootA[E & 2047] * rootB[(E >> 11) & 2047] * rootC[E >> 22]
__device__ u64 d_rootA[2048], d_rootB[2048], d_rootC[1024];
__device__ u32 d_rc3[RC3_WORDS];       // Poseidon round constants as limbs, for per-lane (divergent) indexing

ZKB_D u64 root_pow(u32 E) {
    u64 r = gl_mul(d_rootA[E & 2047], d_rootB[(E >> 11) & 2047]);
    return gl_mul(r, d_rootC[E >> 22]);
}
// w_{2^lg}^e  (e taken mod 2^lg); inverse if inv
ZKB_D u64 root_pow_lg(unsigned lg, u32 e, bool inv) {
    u32 E = lg == 0 ? 0u : (e << (32 - lg));
    if (inv) E = 0u - E;
    return root_pow(E);
}

}  // namespace zkb
#include "ntt.cuh"   // shared-memory NTT kernels (need root_pow)
namespace zkb {

static void ntt_set_func_attributes();   // defined with the NTT kernels below

// The canonical device forms (field.cuh: f_sub, f_add, f_mul, f_canon) put carry-consuming steps on the FMA pipe and rely on the
// polarity of the carry flag a madc sees after sub.cc (1 = no borrow) — compiler behaviour, not a documented contract. Every
// device runs this once before its first use: the forms against the portable compare-and-select arithmetic on corner values and
// a pseudo-random stream. A mismatch makes the library refuse to work on that device instead of producing wrong proofs.
__global__ void field_selfcheck_kernel(unsigned long long* bad) {
    const u64 edge[14] = {0, 1, 2, 7, GL_P - 1, GL_P - 2, 0xffffffffull, 0x100000000ull, 0x100000001ull, 0x8000000000000000ull,
                          GL_P - 0x100000000ull, 0xfffffffe00000002ull, 0x00000001ffffffffull, 0xfffffffeffffffffull};
    const u32 t = threadIdx.x + blockIdx.x * blockDim.x;
    u64 z = 0x9e3779b97f4a7c15ull * (t + 1);
    unsigned long long mism = 0;
    for (int i = 0; i < 256; ++i) {
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z ^= z >> 31;
        const u64 a = i < 196 ? edge[i % 14] : gl_canon(z), b = i < 196 ? edge[(i / 14) % 14] : gl_canon(z * 0x2545f4914f6cdd1dull + t);
        unsigned __int128 pr = (unsigned __int128)a * b;
        const u64 want_mul = (u64)(pr % GL_P);
        mism += f_sub(a, b) != gl_sub(a, b);
        mism += f_add(a, b) != gl_add(a, b);
        mism += f_mul(a, b) != want_mul;
        mism += f_mul(z, b) != (u64)(((unsigned __int128)z * b) % GL_P);          // lazy (non-canonical) left operand
        mism += f_canon(z) != gl_canon(z);
        // the shift forms of the NTT's power-of-two twiddles (every word offset, edge shift counts)
        auto pow2 = [](int k) { u64 r = 1; for (int t = 0; t < k; ++t) r = gl_add(r, r); return r; };
        mism += f_shl<12>(a) != (u64)(((unsigned __int128)a * pow2(12)) % GL_P);
        mism += f_shl<24>(a) != (u64)(((unsigned __int128)a * pow2(24)) % GL_P);
        mism += f_shl<31>(b) != (u64)(((unsigned __int128)b * pow2(31)) % GL_P);
        mism += f_shl<32>(a) != (u64)(((unsigned __int128)a * pow2(32)) % GL_P);
        mism += f_shl<48>(b) != (u64)(((unsigned __int128)b * pow2(48)) % GL_P);
        mism += f_shl<63>(a) != (u64)(((unsigned __int128)a * pow2(63)) % GL_P);
        mism += f_shl<64>(b) != (u64)(((unsigned __int128)b * pow2(64)) % GL_P);
        mism += f_shl<72>(a) != (u64)(((unsigned __int128)a * pow2(72)) % GL_P);
        mism += f_shl<84>(b) != (u64)(((unsigned __int128)b * pow2(84)) % GL_P);
        mism += f_shl<95>(a) != (u64)(((unsigned __int128)a * pow2(95)) % GL_P);
        mism += f_sub_twiddle16<3, true>(a, b) != (u64)(((unsigned __int128)gl_sub(b, a) * pow2(60)) % GL_P);
    }
    if (mism) atomicAdd(bad, mism);
}
static void field_selfcheck() {
    unsigned long long* bad = nullptr;
    unsigned long long h = 0;
    ZKB_CUDA_CHECK(cudaMalloc(&bad, sizeof(h)));
    ZKB_CUDA_CHECK(cudaMemset(bad, 0, sizeof(h)));
    field_selfcheck_kernel<<<16, 128>>>(bad);
    ZKB_CUDA_CHECK(cudaMemcpy(&h, bad, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(bad);
    if (h) throw std::runtime_error("device field arithmetic self-check failed (" + std::to_string(h) +
                                    " mismatches): this build's carry-flag forms do not hold on this toolchain / device");
}

void device_tables_init(int device) {
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lk(mu);
    for (int d : done) if (d == device) return;
    int prev = 0;
    ZKB_CUDA_CHECK(cudaGetDevice(&prev));
    ZKB_CUDA_CHECK(cudaSetDevice(device));
    ZKB_CUDA_CHECK(cudaMemcpyToSymbol(c_rc, host_round_constants(), sizeof(u64) * P_WIDTH * P_ROUNDS));
    ZKB_CUDA_CHECK(cudaMemcpyToSymbol(c_rc3, host_round_constant_limbs(), sizeof(u32) * RC3_WORDS));
    ZKB_CUDA_CHECK(cudaMemcpyToSymbol(d_rc3, host_round_constant_limbs(), sizeof(u32) * RC3_WORDS));
    {
        std::vector<u64> rc2(2 * P_WIDTH * (P_ROUNDS + 1), 0);
        for (int i = 0; i < P_WIDTH * P_ROUNDS; ++i) {
            rc2[2 * i] = host_round_constants()[i] & 0xFFFFFFFFu;
            rc2[2 * i + 1] = host_round_constants()[i] >> 32;
        }
        ZKB_CUDA_CHECK(cudaMemcpyToSymbol(c_rc2, rc2.data(), sizeof(u64) * rc2.size()));
    }
    std::vector<u64> A(2048), B(2048), C(1024);
    u64 T = GL_TWO_ADIC_ROOT, T11 = gl_pow(T, 1u << 11), T22 = gl_pow(T, 1u << 22);
    A[0] = B[0] = C[0] = 1;
    for (int i = 1; i < 2048; ++i) { A[i] = gl_mul(A[i - 1], T); B[i] = gl_mul(B[i - 1], T11); }
    for (int i = 1; i < 1024; ++i) C[i] = gl_mul(C[i - 1], T22);
    ZKB_CUDA_CHECK(cudaMemcpyToSymbol(d_rootA, A.data(), 2048 * 8));
    ZKB_CUDA_CHECK(cudaMemcpyToSymbol(d_rootB, B.data(), 2048 * 8));
    ZKB_CUDA_CHECK(cudaMemcpyToSymbol(d_rootC, C.data(), 1024 * 8));
    const u64 w = gl_root_of_unity(NTT_SM_LG);
    {
        const size_t full = size_t(1) << NTT_SM_LG;
        std::vector<u64> W14(full);
        W14[0] = 1;
        for (size_t i = 1; i < full; ++i) W14[i] = gl_mul(W14[i - 1], w);
        ZKB_CUDA_CHECK(cudaMemcpyToSymbol(d_W14, W14.data(), full * 8));
    }
    ntt_set_func_attributes();
    field_selfcheck();
    ZKB_CUDA_CHECK(cudaSetDevice(prev));
    done.push_back(device);
}

// ---------------------------------------------------------------------------------------------
// Poseidon / Merkle
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) poseidon_permute_kernel(u64* states, size_t count) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) s[k] = states[i * 12 + k];
    poseidon_permute(s);
#pragma unroll
    for (int k = 0; k < 12; ++k) states[i * 12 + k] = gl_canon(s[k]);
}
void launch_poseidon_permute(u64* states, size_t count, cudaStream_t st) {
    if (!count) return;
    ZKB_COUNT_LAUNCH();
    poseidon_permute_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(states, count);
}

ZKB_D void store_digest(u64* digests, size_t idx, u64 d0, u64 d1, u64 d2, u64 d3) {
    ulonglong2* p = reinterpret_cast<ulonglong2*>(digests + idx * 4);
    p[0] = make_ulonglong2(d0, d1);
    p[1] = make_ulonglong2(d2, d3);
}
ZKB_D void store_digest(u64* digests, size_t idx, const PoseidonState& s) {
    store_digest(digests, idx, s.get(0), s.get(1), s.get(2), s.get(3));
}

// one leaf per thread; column-major reads are coalesced across the warp. The sponge state stays in limb form
// between permutations (overwrite mode: the 8 rate words are replaced, the capacity words carry over).
__global__ void __launch_bounds__(128) merkle_leaves_kernel(const u64* __restrict__ leaves, size_t col_stride, int width,
                                                            size_t num_leaves, u64* __restrict__ digests) {
    size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (l >= num_leaves) return;
    if (width <= 4) {   // hash_or_noop: short leaves are copied, not hashed
        u64 d[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < width) d[c] = leaves[(size_t)c * col_stride + l];
        store_digest(digests, l, d[0], d[1], d[2], d[3]);
        return;
    }
    PoseidonState s;
    s.zero();
    const u64* p = leaves + l;
    // one call site of the (large) permutation: a second inlined copy would not fit the instruction cache
#pragma unroll 1
    for (int c = 0; c < width; c += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (c + k < width) s.set(k, __ldg(p + (size_t)(c + k) * col_stride));
        s.permute();
    }
    store_digest(digests, l, s);
}
__global__ void merkle_leaves_wide_kernel(const u64* __restrict__ leaves, size_t col_stride, int width, size_t num_leaves,
                                          u64* __restrict__ digests);   // defined with the lane-parallel kernels below
void launch_merkle_leaves(const u64* leaves, size_t col_stride, int width, size_t num_leaves, u64* digests, cudaStream_t st) {
    if (!num_leaves) return;
    ZKB_COUNT_LAUNCH();
    if (num_leaves <= 4096 && width > 4)   // MERKLE_WIDE_MAX_NODES: a leaf per thread is one chain of ceil(width / 8) permutations
        merkle_leaves_wide_kernel<<<(unsigned)((num_leaves + 7) / 8), 128, 0, st>>>(leaves, col_stride, width, num_leaves, digests);
    else
        merkle_leaves_kernel<<<(unsigned)((num_leaves + 127) / 128), 128, 0, st>>>(leaves, col_stride, width, num_leaves, digests);
}

__global__ void __launch_bounds__(128) merkle_leaves_ext_kernel(const u64* __restrict__ a, const u64* __restrict__ b, int arity,
                                                                size_t num_leaves, u64* __restrict__ digests) {
    size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (l >= num_leaves) return;
    const u64* pa = a + l * arity;
    const u64* pb = b + l * arity;
    int width = 2 * arity;
    if (width <= 4) {
        u64 d[4] = {0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < width) d[c] = (c & 1) ? pb[c >> 1] : pa[c >> 1];
        store_digest(digests, l, d[0], d[1], d[2], d[3]);
        return;
    }
    PoseidonState s;
    s.zero();
#pragma unroll 1
    for (int c = 0; c < width; c += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (c + k < width) s.set(k, (k & 1) ? pb[(c + k) >> 1] : pa[(c + k) >> 1]);
        s.permute();
    }
    store_digest(digests, l, s);
}
__global__ void merkle_leaves_ext_wide_kernel(const u64* __restrict__ a, const u64* __restrict__ b, int arity, size_t num_leaves,
                                              u64* __restrict__ digests);   // defined with the lane-parallel kernels below
void launch_merkle_leaves_ext(const u64* a, const u64* b, int arity, size_t num_leaves, u64* digests, cudaStream_t st) {
    if (!num_leaves) return;
    ZKB_COUNT_LAUNCH();
    if (num_leaves <= 4096 && 2 * arity > 4)   // MERKLE_WIDE_MAX_NODES: latency-bound with one leaf per thread
        merkle_leaves_ext_wide_kernel<<<(unsigned)((num_leaves + 7) / 8), 128, 0, st>>>(a, b, arity, num_leaves, digests);
    else
        merkle_leaves_ext_kernel<<<(unsigned)((num_leaves + 127) / 128), 128, 0, st>>>(a, b, arity, num_leaves, digests);
}

// one parent per thread: parent = permute(left ‖ right ‖ 0000)[0..4]
__global__ void __launch_bounds__(128) merkle_level_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n_out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(in + i * 8);
    ulonglong2 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
    PoseidonState s;
    s.zero();
    s.set(0, v0.x); s.set(1, v0.y); s.set(2, v1.x); s.set(3, v1.y);
    s.set(4, v2.x); s.set(5, v2.y); s.set(6, v3.x); s.set(7, v3.y);
    s.permute();
    store_digest(out, i, s);
}
size_t merkle_level_offset(size_t num_leaves, unsigned level) {
    size_t off = 0;
    for (unsigned k = 0; k < level; ++k) off += num_leaves >> k;
    return off;
}
size_t merkle_digest_count(size_t num_leaves, unsigned cap_height) {
    unsigned lg = 0;
    while ((size_t(1) << lg) < num_leaves) ++lg;
    unsigned levels = lg >= cap_height ? lg - cap_height : 0;
    return merkle_level_offset(num_leaves, levels + 1);
}
// Lane-parallel compression for the SMALL levels of a tree, where one permutation per thread leaves the GPU waiting on
// a single dependency chain (~45 us per level measured): here 12 lanes hold the 12 state words of one node (two nodes
// per warp, lanes 0-11 and 16-27), every lane runs the same S-box code on its own word, and the circulant MDS is
// out_j = sum_i C[i] * x_{(j+i) mod 12}, i.e. 11 warp shuffles per limb. ~4x the issue slots per node, ~1/5 the latency.
// the 30 rounds on one word per lane; jj = this lane's word index (lanes 12-15 of a half-warp shadow word 11)
ZKB_D void wide_permute(u32& x0, u32& x1, u32& x2, const u32* s_rc, unsigned jj) {
#pragma unroll 1
    for (int r = 0; r < P_ROUNDS; ++r) {
        const u32* rc = s_rc + 36 * r + 3 * jj;
        x0 += rc[0]; x1 += rc[1]; x2 += rc[2];
        u32 y0, y1, y2;
        limb_split(gl_sbox7(limb_to_u64(x0, x1, x2)), y0, y1, y2);
        if (!(r < P_HALF_FULL || r >= P_HALF_FULL + P_PARTIAL) && jj != 0) {
            limb_normalize(x0, x1, x2);
            y0 = x0; y1 = x1; y2 = x2;
        }
        x0 = x1 = x2 = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            int src = (int)jj + i;
            if (src >= 12) src -= 12;
            x0 += ZKB_MDS_C(i) * __shfl_sync(0xffffffffu, y0, src, 16);
            x1 += ZKB_MDS_C(i) * __shfl_sync(0xffffffffu, y1, src, 16);
            x2 += ZKB_MDS_C(i) * __shfl_sync(0xffffffffu, y2, src, 16);
        }
        if (jj == 0) { x0 += 8 * y0; x1 += 8 * y1; x2 += 8 * y2; }
    }
}
ZKB_D void wide_load_rc(u32* s_rc) {
    for (unsigned i = threadIdx.x; i < 3 * P_WIDTH * P_ROUNDS; i += blockDim.x) s_rc[i] = d_rc3[i];
    __syncthreads();
}
__global__ void __launch_bounds__(128) merkle_level_wide_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n_out) {
    __shared__ u32 s_rc[3 * P_WIDTH * P_ROUNDS];
    wide_load_rc(s_rc);
    const unsigned lane = threadIdx.x & 31, j = lane & 15, jj = j < 12 ? j : 11;
    const size_t node = ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    const bool live = node < n_out;
    u32 x0, x1, x2;
    limb_split((live && j < 8) ? in[node * 8 + j] : 0, x0, x1, x2);
    wide_permute(x0, x1, x2, s_rc, jj);
    if (live && j < 4) out[node * 4 + j] = gl_canon(limb_to_u64(x0, x1, x2));
}
// lane-parallel sponge over FRI-layer leaves (2 * arity words per leaf, interleaved (a, b) pairs) for the small layers
__global__ void __launch_bounds__(128) merkle_leaves_ext_wide_kernel(const u64* __restrict__ a, const u64* __restrict__ b, int arity,
                                                                     size_t num_leaves, u64* __restrict__ digests) {
    __shared__ u32 s_rc[3 * P_WIDTH * P_ROUNDS];
    wide_load_rc(s_rc);
    const unsigned lane = threadIdx.x & 31, j = lane & 15, jj = j < 12 ? j : 11;
    const size_t l = ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    const bool live = l < num_leaves;
    const int width = 2 * arity;                       // > 4 (the launcher keeps hash_or_noop leaves on the other kernel)
    u32 x0 = 0, x1 = 0, x2 = 0;
#pragma unroll 1
    for (int c = 0; c < width; c += 8) {
        const int k = c + (int)j;
        if (live && j < 8 && k < width) limb_split(((k & 1) ? b : a)[l * arity + (k >> 1)], x0, x1, x2);
        wide_permute(x0, x1, x2, s_rc, jj);
    }
    if (live && j < 4) digests[l * 4 + j] = gl_canon(limb_to_u64(x0, x1, x2));
}
// the same for the row leaves of a polynomial batch (column-major, width > 4): the trees of circuits with n <= 2^9
// (voting-sized) have at most 4096 leaves, far too few threads for one leaf each
__global__ void __launch_bounds__(128) merkle_leaves_wide_kernel(const u64* __restrict__ leaves, size_t col_stride, int width,
                                                                 size_t num_leaves, u64* __restrict__ digests) {
    __shared__ u32 s_rc[3 * P_WIDTH * P_ROUNDS];
    wide_load_rc(s_rc);
    const unsigned lane = threadIdx.x & 31, j = lane & 15, jj = j < 12 ? j : 11;
    const size_t l = ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + (lane >> 4);
    const bool live = l < num_leaves;
    u32 x0 = 0, x1 = 0, x2 = 0;
#pragma unroll 1
    for (int c = 0; c < width; c += 8) {
        const int k = c + (int)j;
        if (live && j < 8 && k < width) limb_split(leaves[(size_t)k * col_stride + l], x0, x1, x2);
        wide_permute(x0, x1, x2, s_rc, jj);
    }
    if (live && j < 4) digests[l * 4 + j] = gl_canon(limb_to_u64(x0, x1, x2));
}

constexpr size_t MERKLE_WIDE_MAX_NODES = 4096;   // levels this small are latency-bound with one node per thread

// ---- tree levels --------------------------------------------------------------------------------------------------------
// One launch per level while a level is real work: one node per thread above 4096 parents (a level of 65 536 nodes is 2.6 % of
// the wires tree's permutations), 12 lanes per node below that (merkle_level_wide_kernel: nodes spread over many CTAs so that
// every SM holds a few lane-parallel chains). The last levels — at most 64 entry nodes per cap digest — are ONE launch:
//   merkle_cap_subtree_kernel  one CTA per cap digest finishes that digest's subtree with the lane-parallel permutation.
// The prover replays each tree's chain of level launches as a CUDA graph (prover.cpp run_merkle_levels).
// Measured and rejected in round 2 (profiles/r02_merkle_fusion.md): (a) building the bottom 7 levels inside the leaf kernel
// (2 launches per tree, bit-exact): every CTA reaches its narrow levels at the same time and holds its registers through seven
// mostly idle permutation latencies — wires tree 2.52 -> 2.78 ms, 8-stream throughput 213 -> 187 proofs/s; (b) one CTA per cap
// digest for the last EIGHT levels (256 entry nodes): 32 warps of lane-parallel chains on one SM serialise on its four
// schedulers (~26 us per pass instead of ~11) — wires tree 2.74 ms.
// levels above `in` (nodes_total digests) down to nodes_total >> levels digests; CTA b owns output digest b of the last level
__global__ void __launch_bounds__(1024) merkle_cap_subtree_kernel(u64* __restrict__ base, size_t nodes_total, int levels) {
    __shared__ u32 s_rc[3 * P_WIDTH * P_ROUNDS];
    wide_load_rc(s_rc);
    const unsigned lane = threadIdx.x & 31, j = lane & 15, jj = j < 12 ? j : 11;
    const unsigned nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5;
    u64* in = base;
    for (int lv = 0; lv < levels; ++lv) {
        const size_t n_in = nodes_total >> lv;
        u64* out = in + n_in * 4;
        const unsigned m_out = (1u << levels) >> (lv + 1);                // parents of this subtree at this level
        const u64* in_sub = in + (size_t)blockIdx.x * 2 * m_out * 4;
        u64* out_sub = out + (size_t)blockIdx.x * m_out * 4;
        for (unsigned b0 = 0; b0 < m_out; b0 += 2 * nwarps) {             // warp-uniform trip count (shuffles inside)
            const unsigned node = b0 + 2 * warp + (lane >> 4);
            const bool live = node < m_out;
            u32 x0, x1, x2;
            limb_split((live && j < 8) ? in_sub[(size_t)node * 8 + j] : 0, x0, x1, x2);
            wide_permute(x0, x1, x2, s_rc, jj);
            if (live && j < 4) out_sub[(size_t)node * 4 + j] = gl_canon(limb_to_u64(x0, x1, x2));
        }
        __syncthreads();                                                  // this CTA's next level reads what it just wrote
        in = out;
    }
}

static unsigned lg2_size(size_t x) { unsigned k = 0; while ((size_t(1) << k) < x) ++k; return k; }
constexpr unsigned MERKLE_CAP_SUBTREE_MAX_ENTRY = 64;      // per cap digest: 32 parents x 16 lanes = one 512-thread CTA per digest

// levels above level `done` (whose digests are in place) of a tree whose leaf level has num_leaves digests; returns the cap offset
static size_t merkle_finish_levels(u64* digests, size_t num_leaves, unsigned cap_height, unsigned done, cudaStream_t st) {
    const unsigned lg = lg2_size(num_leaves);
    const unsigned T = lg >= cap_height ? lg - cap_height : 0;
    unsigned level = done;
    while (level < T) {
        const size_t nodes = num_leaves >> level, off = merkle_level_offset(num_leaves, level);
        const unsigned remaining = T - level;
        if (remaining <= lg2_size(MERKLE_CAP_SUBTREE_MAX_ENTRY) && nodes <= MERKLE_WIDE_MAX_NODES) {
            const unsigned parents = (1u << remaining) >> 1;                // per cap digest
            unsigned threads = parents * 16;
            threads = threads < 32 ? 32 : (threads > 1024 ? 1024 : threads);
            ZKB_COUNT_LAUNCH();
            merkle_cap_subtree_kernel<<<(unsigned)(nodes >> remaining), threads, 0, st>>>(digests + off * 4, nodes, (int)remaining);
            level = T;
        } else {                                                            // one level per launch
            const size_t n_out = nodes >> 1;
            ZKB_COUNT_LAUNCH();
            if (n_out <= MERKLE_WIDE_MAX_NODES)
                merkle_level_wide_kernel<<<(unsigned)((n_out + 7) / 8), 128, 0, st>>>(digests + off * 4, digests + (off + nodes) * 4, n_out);
            else
                merkle_level_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, st>>>(digests + off * 4, digests + (off + nodes) * 4, n_out);
            level += 1;
        }
    }
    return merkle_level_offset(num_leaves, T);
}
size_t launch_merkle_levels(u64* digests, size_t num_leaves, unsigned cap_height, cudaStream_t st) {
    return merkle_finish_levels(digests, num_leaves, cap_height, 0, st);
}
size_t launch_merkle_tree(const u64* leaves, size_t col_stride, int width, size_t num_leaves, u64* digests, unsigned cap_height,
                          cudaStream_t st) {
    if (!num_leaves) return 0;
    launch_merkle_leaves(leaves, col_stride, width, num_leaves, digests, st);
    return merkle_finish_levels(digests, num_leaves, cap_height, 0, st);
}
size_t launch_merkle_tree_ext(const u64* a_, const u64* b_, int arity, size_t num_leaves, u64* digests, unsigned cap_height,
                              cudaStream_t st) {
    if (!num_leaves) return 0;
    launch_merkle_leaves_ext(a_, b_, arity, num_leaves, digests, st);
    return merkle_finish_levels(digests, num_leaves, cap_height, 0, st);
}

// every element must be a canonical field element (< p): *flag |= 1 otherwise (the reference's types guarantee this; a
// non-canonical u64 from a foreign caller would otherwise give a silently wrong proof)
__global__ void canonical_check_kernel(const u64* __restrict__ v, size_t count, unsigned* flag) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < count; i += stride) bad |= v[i] >= GL_P;
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}
void launch_canonical_check(const u64* v, size_t count, unsigned* flag_dev, cudaStream_t st) {
    if (!count) return;
    size_t blocks = (count + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    ZKB_COUNT_LAUNCH();
    canonical_check_kernel<<<(unsigned)blocks, 256, 0, st>>>(v, count, flag_dev);
}

// ---------------------------------------------------------------------------------------------
// salts (documented generator shared with the oracle: oracle/prover.hpp salt_value)
// ---------------------------------------------------------------------------------------------
ZKB_HD u64 splitmix64_finalize(u64 z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__global__ void salt_fill_kernel(u64* out, size_t stride, size_t num_leaves, u64 seed, unsigned batch) {
    size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    unsigned s = blockIdx.y;
    if (l >= num_leaves) return;
    u64 z = seed + 0x9e3779b97f4a7c15ULL * (u64)(batch * 4 + s + 1) + l;
    out[(size_t)s * stride + l] = gl_canon(splitmix64_finalize(z + 0x9e3779b97f4a7c15ULL));
}
void launch_salt_fill(u64* out, size_t stride, size_t num_leaves, u64 seed, unsigned batch, cudaStream_t st) {
    dim3 grid((unsigned)((num_leaves + 255) / 256), 4);
    ZKB_COUNT_LAUNCH();
    salt_fill_kernel<<<grid, 256, 0, st>>>(out, stride, num_leaves, seed, batch);
}

// Production salts: the blinding columns hide the opened leaves' neighbours only if they are unpredictable, so they come from
// a CSPRNG — ChaCha20 (RFC 8439 block function, 20 rounds) keyed per proof with 256 bits the host reads from the OS
// (getrandom). Block t of the key stream (64-bit counter t, nonce word = batch) gives 8 consecutive words of the flat
// [4][stride] salt array; a 64-bit word is mapped to the field by one conditional subtraction of p (statistical distance from
// uniform 2^-32 per element — qp-plonky2 samples F::rand by rejection; the difference is not observable by a verifier).
struct ChaChaKey { u32 k[8]; };
ZKB_D u32 rotl32(u32 x, int r) { return __funnelshift_l(x, x, r); }
#define ZKB_QR(a, b, c, d) a += b; d = rotl32(d ^ a, 16); c += d; b = rotl32(b ^ c, 12); a += b; d = rotl32(d ^ a, 8); c += d; b = rotl32(b ^ c, 7)
__global__ void __launch_bounds__(256) salt_chacha_kernel(u64* out, size_t words, ChaChaKey key, u32 batch) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t * 8 >= words) return;
    u32 in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5],
                  key.k[6], key.k[7], (u32)t, (u32)(t >> 32), batch, 0x7a6b6232u};
    u32 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = in[i];
#pragma unroll 2
    for (int r = 0; r < 10; ++r) {
        ZKB_QR(x[0], x[4], x[8], x[12]); ZKB_QR(x[1], x[5], x[9], x[13]); ZKB_QR(x[2], x[6], x[10], x[14]); ZKB_QR(x[3], x[7], x[11], x[15]);
        ZKB_QR(x[0], x[5], x[10], x[15]); ZKB_QR(x[1], x[6], x[11], x[12]); ZKB_QR(x[2], x[7], x[8], x[13]); ZKB_QR(x[3], x[4], x[9], x[14]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const u64 v = gl_pack(x[2 * i] + in[2 * i], x[2 * i + 1] + in[2 * i + 1]);
        if (t * 8 + i < words) out[t * 8 + i] = gl_canon(v);
    }
}
void launch_salt_fill_csprng(u64* out, size_t words, const u32 key[8], unsigned batch, cudaStream_t st) {
    if (!words) return;
    ChaChaKey k;
    for (int i = 0; i < 8; ++i) k.k[i] = key[i];
    const size_t blocks = (words / 8 + 255) / 256 + 1;
    ZKB_COUNT_LAUNCH();
    salt_chacha_kernel<<<(unsigned)blocks, 256, 0, st>>>(out, words, k, (u32)batch);
}

// ---------------------------------------------------------------------------------------------
// NTT family: host side. Kernels are in ntt.cuh. n <= 2^14: one shared-memory kernel per transform; larger n: two steps
// (n = n1 * n2: n1-point transforms down TB-wide column tiles + twiddle, then contiguous n2-point transforms).
// ---------------------------------------------------------------------------------------------
// n <= 2^20: the first step runs in registers (ntt_cols_reg_kernel: n1 <= 64 points per thread, shift twiddles, one merged table
// multiply), the second is the compile-time 2^14 block kernel. ZKB_NTT_COLS_GENERIC=1 (or n > 2^20) selects the shared-memory
// first step.
static bool large_in_registers() {
    static const bool on = [] { const char* e = std::getenv("ZKB_NTT_COLS_GENERIC"); return !(e && e[0] == '1'); }();
    return on;
}
// in registers: n = 2^lg_n1 x 2^lg_mid x 2^lb with lg_n1 <= 6; lg_mid = 6 for n > 2^20 (three steps: the rows the first step
// leaves are plain 2^20-point transforms), 0 otherwise
struct LargePlan { unsigned lb, lg_n1, lg_tb, lg_mid; };
static LargePlan large_plan(unsigned lg_n) {
    if (lg_n > NTT_SM_LG + 10) throw std::runtime_error("NTT size too large (max 2^24)");
    LargePlan p;
    p.lg_mid = 0;
    if (large_in_registers()) {
        p.lb = NTT_SM_LG;
        p.lg_mid = lg_n > NTT_SM_LG + 6 ? 6 : 0;
        p.lg_n1 = lg_n - NTT_SM_LG - p.lg_mid;
        p.lg_tb = 0;
        return p;
    }
    p.lb = lg_n >= 20 ? (lg_n - 8 > NTT_SM_LG ? NTT_SM_LG : (lg_n - 8 < 12 ? 12 : lg_n - 8)) : 12;
    if (lg_n - p.lb > 10) p.lb = lg_n - 10;
    p.lg_n1 = lg_n - p.lb;
    p.lg_tb = p.lg_n1 >= 8 ? 4 : 12 - p.lg_n1;          // tile = n1 x TB >= 4096 elements, >= 128 B per row segment
    return p;
}
static void ntt_set_func_attributes() {   // per device: opt in to > 48 KB dynamic shared memory
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(lde_block_kernel_t<4, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(NTT_SM_LG)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(lde_block_kernel_t<3, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(NTT_SM_LG)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(lde_block_kernel_t<3, 1024, 14>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(14)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(lde_block_kernel_t<3, 1024, 13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(13)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(lde_block_kernel_t<3, 512, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(12)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(intt_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(NTT_SM_LG)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(intt_block_kernel_c<14, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(14)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(intt_block_kernel_c<13, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(13)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(intt_block_kernel_c<12, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(12)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(ntt_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ntt_smem_bytes(NTT_SM_LG)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(ntt_cols_staged_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(u64) * 128 * 64)));
    ZKB_CUDA_CHECK(cudaFuncSetAttribute(ntt_cols_staged_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(u64) * 128 * 64)));
}
static unsigned ntt_block_threads(unsigned lg_n) {
    unsigned t = lg_n >= 4 ? (1u << (lg_n - 4)) : 1u;
    return t < 32 ? 32 : (t > 512 ? 512 : t);
}
// Blocks of n >= 2^13 run radix-8 passes with 1024 threads (64 registers, 32 warps per SM) — measured 5 % faster than
// radix-16 passes with 512 threads (128 registers) on the 135-column wormhole batch (profiles/r01_bench_v6_ntt_variants.json:
// 0.368 vs 0.389 ms); ZKB_NTT_RADIX8=0 selects the radix-16 form.
static bool ntt_radix8() {
    static const bool on = [] { const char* e = std::getenv("ZKB_NTT_RADIX8"); return !(e && e[0] == '0'); }();
    return on;
}
static bool ntt_specialised() {
    static const bool on = [] { const char* e = std::getenv("ZKB_NTT_GENERIC"); return !(e && e[0] == '1'); }();
    return on;
}
static void launch_lde_block(dim3 grid, unsigned lg_n, cudaStream_t st, const u64* coeffs, size_t coeff_stride, u64* out,
                             size_t out_stride, const u64* prescale, size_t src_block_stride, int inv, unsigned jb0) {
    ZKB_COUNT_LAUNCH();
    const bool specialised = ntt_specialised();
    if (specialised && ntt_radix8() && lg_n == 14)
        lde_block_kernel_t<3, 1024, 14><<<grid, 1024, ntt_smem_bytes(14), st>>>(coeffs, coeff_stride, out, out_stride, lg_n, prescale,
                                                                               src_block_stride, inv, jb0);
    else if (specialised && ntt_radix8() && lg_n == 13)
        lde_block_kernel_t<3, 1024, 13><<<grid, 1024, ntt_smem_bytes(13), st>>>(coeffs, coeff_stride, out, out_stride, lg_n, prescale,
                                                                               src_block_stride, inv, jb0);
    else if (specialised && ntt_radix8() && lg_n == 12)
        lde_block_kernel_t<3, 512, 12><<<grid, 512, ntt_smem_bytes(12), st>>>(coeffs, coeff_stride, out, out_stride, lg_n, prescale,
                                                                             src_block_stride, inv, jb0);
    else if (lg_n >= 13 && ntt_radix8())
        lde_block_kernel_t<3, 1024><<<grid, 1024, ntt_smem_bytes(lg_n), st>>>(coeffs, coeff_stride, out, out_stride, lg_n, prescale,
                                                                             src_block_stride, inv, jb0);
    else
        lde_block_kernel_t<4, 512><<<grid, ntt_block_threads(lg_n), ntt_smem_bytes(lg_n), st>>>(coeffs, coeff_stride, out, out_stride,
                                                                                               lg_n, prescale, src_block_stride, inv, jb0);
}
// Coset pre-scale tables (shift * w_N^j)^k, cached per device for the lifetime of the process. n <= 2^14: one table
// [2^rate][n] (1 MB for the wormhole circuit). Larger n: the exponent is split k = r * n2 + b, tables [2^rate][n1], [2^rate][n2].
struct CosetTables { u64* full = nullptr; u64* pre1 = nullptr; u64* pre2 = nullptr; u64* tw = nullptr; };
__global__ void coset_table2_kernel(u64* pre1, u64* pre2, unsigned lg_n1, unsigned lg_n2, unsigned rate_bits, u64 shift, u64 w_N) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x, jb = blockIdx.y;
    const u64 base = gl_mul(shift, gl_pow(w_N, bitrev32(jb, rate_bits)));
    if (pre2 && k < (1u << lg_n2)) pre2[((size_t)jb << lg_n2) + k] = gl_pow(base, k);
    if (k < (1u << lg_n1)) pre1[((size_t)jb << lg_n1) + k] = gl_pow(gl_pow(base, u64(1) << lg_n2), k);
}
static CosetTables coset_tables(unsigned lg_n, unsigned rate_bits, u64 shift, cudaStream_t st) {
    static std::mutex mu;
    static std::map<std::tuple<int, unsigned, unsigned, u64>, CosetTables> cache;
    int dev = 0;
    ZKB_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(dev, lg_n, rate_bits, shift);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    CosetTables t;
    const u64 w_N = gl_root_of_unity(lg_n + rate_bits);
    if (lg_n <= NTT_SM_LG) {
        const size_t n = size_t(1) << lg_n;
        ZKB_CUDA_CHECK(cudaMalloc(&t.full, (n << rate_bits) * sizeof(u64)));
        dim3 grid((unsigned)((n + 127) / 128), 1u << rate_bits);
        ZKB_COUNT_LAUNCH();
        coset_table_kernel<<<grid, 128, 0, st>>>(t.full, lg_n, rate_bits, shift, w_N);
    } else {
        const LargePlan p = large_plan(lg_n);
        ZKB_CUDA_CHECK(cudaMalloc(&t.pre1, (sizeof(u64) << p.lg_n1) << rate_bits));
        if (large_in_registers()) {            // merged s_j^b w_n^(b bitrev(p)) table: one LDE column's worth of memory per device
            const unsigned lg_row = lg_n - p.lg_n1;
            dim3 grid(((1u << p.lg_n1) + 127) / 128, 1u << rate_bits);
            ZKB_COUNT_LAUNCH();
            coset_table2_kernel<<<grid, 128, 0, st>>>(t.pre1, nullptr, p.lg_n1, lg_row, rate_bits, shift, w_N);
            ZKB_CUDA_CHECK(cudaMalloc(&t.tw, (sizeof(u64) << lg_n) << rate_bits));
            dim3 g((1u << lg_row) / 128, 1u << p.lg_n1, 1u << rate_bits);
            ZKB_COUNT_LAUNCH();
            cols_reg_table_kernel<<<g, 128, 0, st>>>(t.tw, p.lg_n1, lg_row, rate_bits, shift, w_N, gl_root_of_unity(lg_n), 1);
        } else {
            ZKB_CUDA_CHECK(cudaMalloc(&t.pre2, (sizeof(u64) << p.lb) << rate_bits));
            const unsigned m = 1u << (p.lb > p.lg_n1 ? p.lb : p.lg_n1);
            dim3 grid((m + 127) / 128, 1u << rate_bits);
            ZKB_COUNT_LAUNCH();
            coset_table2_kernel<<<grid, 128, 0, st>>>(t.pre1, t.pre2, p.lg_n1, p.lb, rate_bits, shift, w_N);
        }
    }
    ZKB_CUDA_CHECK(cudaStreamSynchronize(st));       // other streams may use the tables as soon as they are in the cache
    cache[key] = t;
    return t;
}
// w_n^(+-b bitrev(p)), p < 2^lg_n1, b < n / 2^lg_n1, for the plain (no coset) in-register step, cached per device
static const u64* plain_cols_table(unsigned lg_n, unsigned lg_n1, bool inv, cudaStream_t st) {
    static std::mutex mu;
    static std::map<std::tuple<int, unsigned, unsigned, bool>, u64*> cache;
    int dev = 0;
    ZKB_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(dev, lg_n, lg_n1, inv);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    u64* tw = nullptr;
    ZKB_CUDA_CHECK(cudaMalloc(&tw, sizeof(u64) << lg_n));
    u64 w = gl_root_of_unity(lg_n);
    if (inv) w = gl_inv(w);
    dim3 g((1u << (lg_n - lg_n1)) / 128, 1u << lg_n1, 1);
    ZKB_COUNT_LAUNCH();
    cols_reg_table_kernel<<<g, 128, 0, st>>>(tw, lg_n1, lg_n - lg_n1, 0, 1, 1, w, 0);
    ZKB_CUDA_CHECK(cudaStreamSynchronize(st));
    cache[key] = tw;
    return tw;
}
template <bool INV>
static void launch_cols_reg(dim3 grid, unsigned lg_n1, const ColsRegArgs& a, cudaStream_t st) {
    ZKB_COUNT_LAUNCH();
    switch (lg_n1) {
        case 1: ntt_cols_reg_kernel<1, INV><<<grid, 128, 0, st>>>(a); break;
        case 2: ntt_cols_reg_kernel<2, INV><<<grid, 128, 0, st>>>(a); break;
        case 3: ntt_cols_reg_kernel<3, INV><<<grid, 128, 0, st>>>(a); break;
        case 4: ntt_cols_reg_kernel<4, INV><<<grid, 128, 0, st>>>(a); break;
        case 5: ntt_cols_staged_kernel<5, INV><<<grid, 128, sizeof(u64) * 128 * 32, st>>>(a); break;
        default: ntt_cols_staged_kernel<6, INV><<<grid, 128, sizeof(u64) * 128 * 64, st>>>(a); break;
    }
}
// the two steps for n > 2^14; z = number of coset blocks written (block jb of dst at + jb * n). tw: merged coset table of the
// in-register first step (null: plain transform)
static void run_large_transform(const u64* src, size_t src_stride, u64* dst, size_t dst_stride, int ncols, unsigned lg_n,
                                unsigned nblk, const u64* pre1, const u64* pre2, bool inv, cudaStream_t st, unsigned jb0 = 0,
                                const u64* tw = nullptr) {
    const LargePlan p = large_plan(lg_n);
    if (large_in_registers()) {
        const unsigned lg_row = lg_n - p.lg_n1;         // 14, or 20 when a middle step follows
        ColsRegArgs a{src, src_stride, dst, dst_stride, lg_row, nblk, jb0, pre1, tw, size_t(1) << lg_n, 1, 0};
        if (!pre1) { a.tw = plain_cols_table(lg_n, p.lg_n1, inv, st); a.tw_block_stride = 0; a.jb0 = 0; }
        dim3 g1((unsigned)ncols, (1u << lg_row) / 128);
        if (inv) launch_cols_reg<true>(g1, p.lg_n1, a, st);
        else launch_cols_reg<false>(g1, p.lg_n1, a, st);
        if (p.lg_mid) {                                 // the nblk * n1 rows of a column: plain 2^20-point transforms, in place
            const unsigned nsub = nblk << p.lg_n1;
            ColsRegArgs m{dst, dst_stride, dst, dst_stride, p.lb, 1, 0, nullptr, plain_cols_table(lg_row, p.lg_mid, inv, st), 0,
                          nsub, size_t(1) << lg_row};
            dim3 gm((unsigned)ncols * nsub, (1u << p.lb) / 128);
            if (inv) launch_cols_reg<true>(gm, p.lg_mid, m, st);
            else launch_cols_reg<false>(gm, p.lg_mid, m, st);
        }
        dim3 g2(nblk << (lg_n - p.lb), (unsigned)ncols);
        launch_lde_block(g2, p.lb, st, dst, dst_stride, dst, dst_stride, nullptr, size_t(1) << p.lb, inv ? 1 : 0, 0);
        return;
    }
    ColsNttArgs a{src, src_stride, dst, dst_stride, lg_n, p.lg_n1, p.lg_tb, pre1, pre2, inv ? 1 : 0, jb0};
    dim3 g1(1u << (p.lb - p.lg_tb), (unsigned)ncols, nblk);
    ZKB_COUNT_LAUNCH();
    ntt_cols_kernel<<<g1, 256, ntt_smem_bytes(p.lg_n1 + p.lg_tb), st>>>(a);
    dim3 g2(nblk << p.lg_n1, (unsigned)ncols);
    launch_lde_block(g2, p.lb, st, dst, dst_stride, dst, dst_stride, nullptr, size_t(1) << p.lb, inv ? 1 : 0, 0);
}

// in-place bit-reversal permutation with scaling, n >= 2^11: index = (hi : mid : lo) with 5-bit hi and lo; a CTA swaps the
// 32 x 32 tiles of mid and bitrev(mid) through shared memory, so that reads and writes are both 256-byte row segments (the
// element-wise kernel below scatters one side: 1.0 TB/s on 2^20-point columns)
__global__ void __launch_bounds__(256) bitrev_scale_tiled_kernel(u64* data, size_t stride, unsigned lg_n, u64 scale) {
    __shared__ u64 ta[32][33], tb[32][33];
    const unsigned lg_mid = lg_n - 10, mid = blockIdx.x, rmid = bitrev32(mid, lg_mid);
    if (mid > rmid) return;
    u64* col = data + (size_t)blockIdx.y * stride;
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned sh_hi = lg_n - 5;
#pragma unroll
    for (unsigned hi = w; hi < 32; hi += 8) {
        ta[hi][lane] = col[((size_t)hi << sh_hi) | ((size_t)mid << 5) | lane];
        if (mid != rmid) tb[hi][lane] = col[((size_t)hi << sh_hi) | ((size_t)rmid << 5) | lane];
    }
    __syncthreads();
    const unsigned rl = __brev(lane) >> 27;
#pragma unroll
    for (unsigned hi = w; hi < 32; hi += 8) {           // new[hi][bitrev(mid)][lane] = old[bitrev(lane)][mid][bitrev(hi)]
        const unsigned rh = __brev(hi) >> 27;
        col[((size_t)hi << sh_hi) | ((size_t)rmid << 5) | lane] = gl_mul(ta[rl][rh], scale);
        if (mid != rmid) col[((size_t)hi << sh_hi) | ((size_t)mid << 5) | lane] = gl_mul(tb[rl][rh], scale);
    }
}
__global__ void bitrev_scale_kernel(u64* data, size_t stride, unsigned lg_n, u64 scale) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= (size_t(1) << lg_n)) return;
    u64* col = data + (size_t)blockIdx.y * stride;
    size_t r = bitrev32((u32)k, lg_n);
    if (k < r) {
        u64 x = col[k], y = col[r];
        col[k] = gl_mul(y, scale);
        col[r] = gl_mul(x, scale);
    } else if (k == r) {
        col[k] = gl_mul(col[k], scale);
    }
}
static void run_bitrev_scale(u64* data, size_t stride, int ncols, unsigned lg_n, u64 scale, cudaStream_t st) {
    ZKB_COUNT_LAUNCH();
    if (lg_n >= 11) {
        dim3 tiles(1u << (lg_n - 10), (unsigned)ncols);
        bitrev_scale_tiled_kernel<<<tiles, 256, 0, st>>>(data, stride, lg_n, scale);
        return;
    }
    dim3 grid((unsigned)(((size_t(1) << lg_n) + 255) / 256), (unsigned)ncols);
    bitrev_scale_kernel<<<grid, 256, 0, st>>>(data, stride, lg_n, scale);
}
void launch_bitrev_permute(u64* data, size_t stride, int ncols, unsigned lg_n, cudaStream_t st) {
    if (ncols <= 0) return;
    run_bitrev_scale(data, stride, ncols, lg_n, 1, st);
}

// one column per CTA: values -> coefficients (times scale), specialised for the circuit sizes
static void launch_intt_block(const u64* in, size_t in_stride, u64* out, size_t out_stride, int ncols, unsigned lg_n, int in_bitrev,
                              u64 scale, cudaStream_t st) {
    ZKB_COUNT_LAUNCH();
    if (ntt_specialised() && lg_n == 14)
        intt_block_kernel_c<14, 1024><<<(unsigned)ncols, 1024, ntt_smem_bytes(14), st>>>(in, in_stride, out, out_stride, in_bitrev, scale);
    else if (ntt_specialised() && lg_n == 13)
        intt_block_kernel_c<13, 1024><<<(unsigned)ncols, 1024, ntt_smem_bytes(13), st>>>(in, in_stride, out, out_stride, in_bitrev, scale);
    else if (ntt_specialised() && lg_n == 12)
        intt_block_kernel_c<12, 512><<<(unsigned)ncols, 512, ntt_smem_bytes(12), st>>>(in, in_stride, out, out_stride, in_bitrev, scale);
    else
        intt_block_kernel<<<(unsigned)ncols, ntt_block_threads(lg_n), ntt_smem_bytes(lg_n), st>>>(in, in_stride, out, out_stride, lg_n, in_bitrev, scale);
}
void launch_intt_natural(const u64* src, size_t src_stride, u64* dst, size_t dst_stride, int ncols, unsigned lg_n,
                         u64* scratch, cudaStream_t st) {
    (void)scratch;
    if (ncols <= 0) return;
    u64 ninv = gl_inv(u64(1) << lg_n);
    if (lg_n <= NTT_SM_LG) {
        launch_intt_block(src, src_stride, dst, dst_stride, ncols, lg_n, 0, ninv, st);
        return;
    }
    run_large_transform(src, src_stride, dst, dst_stride, ncols, lg_n, 1, nullptr, nullptr, true, st);   // bit-reversed result
    run_bitrev_scale(dst, dst_stride, ncols, lg_n, ninv, st);
}

void launch_lde_blocks(const u64* coeffs, size_t coeff_stride, u64* out, size_t out_stride, int ncols, unsigned lg_n,
                       unsigned rate_bits, u64 shift, unsigned blk_lo, unsigned blk_hi, cudaStream_t st) {
    if (ncols <= 0 || blk_hi <= blk_lo) return;
    const bool plain = rate_bits == 0 && shift == 1;
    CosetTables t;
    if (!plain) t = coset_tables(lg_n, rate_bits, shift, st);
    if (lg_n <= NTT_SM_LG) {
        dim3 grid(blk_hi - blk_lo, (unsigned)ncols);
        launch_lde_block(grid, lg_n, st, coeffs, coeff_stride, out, out_stride, t.full, 0, 0, blk_lo);
        return;
    }
    run_large_transform(coeffs, coeff_stride, out, out_stride, ncols, lg_n, blk_hi - blk_lo, t.pre1, t.pre2, false, st, blk_lo, t.tw);
}
void launch_lde(const u64* coeffs, size_t coeff_stride, u64* out, size_t out_stride, int ncols, unsigned lg_n,
                unsigned rate_bits, u64 shift, cudaStream_t st) {
    launch_lde_blocks(coeffs, coeff_stride, out, out_stride, ncols, lg_n, rate_bits, shift, 0, 1u << rate_bits, st);
}

// out[m][k] = sum_j mat[m][j] in[j][k]: recovers the R chunk polynomials from the R per-coset interpolants (sharded.hpp)
__global__ void vandermonde_solve_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t len, unsigned R, const u64* __restrict__ mat) {
    const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= len) return;
    u64 v[16];
    for (unsigned j = 0; j < R; ++j) v[j] = in[(size_t)j * len + k];
    for (unsigned m = 0; m < R; ++m) {
        u64 acc = 0;
        for (unsigned j = 0; j < R; ++j) acc = f_add(acc, f_mul(v[j], mat[m * R + j]));
        out[(size_t)m * len + k] = acc;
    }
}
void launch_vandermonde_solve(const u64* in, u64* out, size_t len, unsigned R, const u64* mat_dev, cudaStream_t st) {
    if (!len) return;
    ZKB_COUNT_LAUNCH();
    vandermonde_solve_kernel<<<(unsigned)((len + 127) / 128), 128, 0, st>>>(in, out, len, R, mat_dev);
}

// data[k] *= c0 * base^k
__global__ void scale_pows_kernel(u64* data, size_t stride, int ncols, size_t m, u64 c0, u64 base) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= m) return;
    u64 f = gl_mul(c0, gl_pow(base, k));
    for (int c = 0; c < ncols; ++c) data[(size_t)c * stride + k] = gl_mul(data[(size_t)c * stride + k], f);
}

void launch_coset_intt_bitrev(u64* data, size_t stride, int ncols, unsigned lg_m, u64 shift, cudaStream_t st) {
    if (ncols <= 0) return;
    size_t m = size_t(1) << lg_m;
    if (lg_m <= NTT_SM_LG) {
        launch_intt_block(data, stride, data, stride, ncols, lg_m, 1, 1, st);
    } else {
        run_bitrev_scale(data, stride, ncols, lg_m, 1, st);                                             // leaf order -> natural
        run_large_transform(data, stride, data, stride, ncols, lg_m, 1, nullptr, nullptr, true, st);   // -> bit-reversed coefficients
        run_bitrev_scale(data, stride, ncols, lg_m, 1, st);
    }
    ZKB_COUNT_LAUNCH();
    scale_pows_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(data, stride, ncols, m, gl_inv(u64(1) << lg_m), gl_inv(shift));
}


// ---------------------------------------------------------------------------------------------
// Partial products and Z: per-row chunk quotients, exact parallel prefix product
// ---------------------------------------------------------------------------------------------
struct PPArgs {
    const u64* wires; size_t wire_stride;
    const u64* sigmas; size_t sigma_stride;
    const u64* k_is;
    int num_routed, chunk, nchunks, nch;
    unsigned lg_n;
    u64 betas[4], gammas[4];
    u64* chunkprod;   // [nch][nchunks][n]
    u64* rowtot;      // [nch][n]
    u64* prefix;      // [nch][n]
    u64* out; size_t out_stride;
};

constexpr int PP_MAX_CHUNKS = 16;
// one copy of the 84-multiplication inversion chain instead of one per unrolled chunk slot (the kernel was 310 KB of code)
__device__ __noinline__ u64 gl_inv_call(u64 a) { return gl_inv(a); }
__global__ void __launch_bounds__(128) pp_chunk_kernel(PPArgs a) {
    size_t n = size_t(1) << a.lg_n;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    int ch = blockIdx.y;
    if (i >= n) return;
    u64 beta = a.betas[ch], gamma = a.gammas[ch];
    u64 bx = gl_mul(beta, root_pow_lg(a.lg_n, (u32)i, false));
    // numerators, and PREFIX PRODUCTS of the denominators: one inversion per row instead of one per chunk
    u64 num[PP_MAX_CHUNKS], den[PP_MAX_CHUNKS], pre[PP_MAX_CHUNKS];
    u64 run = 1;
#pragma unroll
    for (int k = 0; k < PP_MAX_CHUNKS; ++k) {
        num[k] = den[k] = pre[k] = 1;
        if (k < a.nchunks) {
            u64 nu = 1, de = 1;
            int j1 = min(a.num_routed, (k + 1) * a.chunk);
#pragma unroll 1
            for (int j = k * a.chunk; j < j1; ++j) {
                u64 w = a.wires[(size_t)j * a.wire_stride + i];
                u64 wg = f_add(w, gamma);
                nu = f_mul(nu, f_add(wg, f_mul(bx, a.k_is[j])));
                de = f_mul(de, f_add(wg, f_mul(beta, a.sigmas[(size_t)j * a.sigma_stride + i])));
            }
            num[k] = nu; den[k] = de;
            pre[k] = run;                       // product of den[0..k)
            run = gl_mul(run, de);
        }
    }
    const bool degenerate = run == 0;            // some denominator is zero: keep the per-chunk definition (inv(0) = 0)
    u64 inv_run = gl_inv_call(run);
    u64 tot = 1;
    u64 q[PP_MAX_CHUNKS];
#pragma unroll
    for (int k = PP_MAX_CHUNKS - 1; k >= 0; --k) {
        q[k] = 1;
        if (k < a.nchunks) {
            u64 inv_k = degenerate ? gl_inv_call(den[k]) : gl_mul(inv_run, pre[k]);
            inv_run = gl_mul(inv_run, den[k]);
            q[k] = gl_mul(num[k], inv_k);
        }
    }
#pragma unroll
    for (int k = 0; k < PP_MAX_CHUNKS; ++k) {
        if (k < a.nchunks) {
            a.chunkprod[((size_t)ch * a.nchunks + k) * n + i] = q[k];
            tot = gl_mul(tot, q[k]);
        }
    }
    a.rowtot[(size_t)ch * n + i] = tot;
}

// exclusive prefix product over n rows, one CTA per challenge
__global__ void __launch_bounds__(1024) pp_scan_kernel(PPArgs a) {
    __shared__ u64 part[1024];
    size_t n = size_t(1) << a.lg_n;
    int ch = blockIdx.x;
    const u64* in = a.rowtot + (size_t)ch * n;
    u64* out = a.prefix + (size_t)ch * n;
    unsigned nt = blockDim.x;
    size_t ipt = n / nt;                 // launcher guarantees nt | n
    size_t lo = threadIdx.x * ipt;
    u64 p = 1;
    for (size_t k = 0; k < ipt; ++k) p = gl_mul(p, in[lo + k]);
    part[threadIdx.x] = p;
    __syncthreads();
    for (unsigned d = 1; d < nt; d <<= 1) {      // Hillis-Steele inclusive scan (exact: field mult is associative)
        u64 v = part[threadIdx.x];
        if (threadIdx.x >= d) v = gl_mul(v, part[threadIdx.x - d]);
        __syncthreads();
        part[threadIdx.x] = v;
        __syncthreads();
    }
    u64 acc = threadIdx.x ? part[threadIdx.x - 1] : 1;
    for (size_t k = 0; k < ipt; ++k) {
        out[lo + k] = acc;
        acc = gl_mul(acc, in[lo + k]);
    }
}

__global__ void __launch_bounds__(128) pp_finalize_kernel(PPArgs a) {
    size_t n = size_t(1) << a.lg_n;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    int ch = blockIdx.y;
    if (i >= n) return;
    u64 z = a.prefix[(size_t)ch * n + i];
    a.out[(size_t)ch * a.out_stride + i] = z;
    int npp = a.nchunks - 1;
    u64 acc = z;
    for (int k = 0; k < npp; ++k) {
        acc = gl_mul(acc, a.chunkprod[((size_t)ch * a.nchunks + k) * n + i]);
        a.out[(size_t)(a.nch + ch * npp + k) * a.out_stride + i] = acc;
    }
}

size_t partial_products_scratch_words(int num_routed, int chunk, int num_challenges, unsigned lg_n) {
    size_t nchunks = (num_routed + chunk - 1) / chunk;
    return (size_t)num_challenges * (nchunks + 2) << lg_n;
}
void launch_partial_products(const u64* wires, size_t wire_stride, const u64* sigma_values, size_t sigma_stride,
                             const u64* k_is_dev, int num_routed, int chunk, int num_challenges, const u64* betas_gammas,
                             unsigned lg_n, u64* out, size_t out_stride, u64* scratch, cudaStream_t st) {
    PPArgs a;
    a.wires = wires; a.wire_stride = wire_stride; a.sigmas = sigma_values; a.sigma_stride = sigma_stride;
    a.k_is = k_is_dev; a.num_routed = num_routed; a.chunk = chunk; a.nchunks = (num_routed + chunk - 1) / chunk;
    a.nch = num_challenges; a.lg_n = lg_n;
    if (a.nchunks > PP_MAX_CHUNKS) throw std::runtime_error("too many partial-product chunks");
    for (int c = 0; c < num_challenges; ++c) { a.betas[c] = betas_gammas[c]; a.gammas[c] = betas_gammas[num_challenges + c]; }
    size_t n = size_t(1) << lg_n;
    a.chunkprod = scratch;
    a.rowtot = scratch + (size_t)num_challenges * a.nchunks * n;
    a.prefix = a.rowtot + (size_t)num_challenges * n;
    a.out = out; a.out_stride = out_stride;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)num_challenges);
    ZKB_COUNT_LAUNCH();
    pp_chunk_kernel<<<grid, 128, 0, st>>>(a);
    unsigned nt = n < 1024 ? (unsigned)n : 1024;
    ZKB_COUNT_LAUNCH();
    pp_scan_kernel<<<num_challenges, nt, 0, st>>>(a);
    ZKB_COUNT_LAUNCH();
    pp_finalize_kernel<<<grid, 128, 0, st>>>(a);
}

// ---------------------------------------------------------------------------------------------
// Quotient: vanishing polynomial at every LDE point / Z_H, gate set interpreted from
// the circuit's gate list (warp-uniform dispatch). Gate formulas: SURVEY App. C.1.
// ---------------------------------------------------------------------------------------------
constexpr u32 TAG_ARITHMETIC = 0, TAG_BASE_SUM = 2, TAG_CONSTANT = 3, TAG_NOOP = 9, TAG_POSEIDON = 11, TAG_PUBLIC_INPUT = 12;
// recursion gate set, evaluated by the third launch
constexpr u32 TAG_ARITHMETIC_EXT = 1, TAG_COSET_INTERP = 4, TAG_EXPONENTIATION = 5, TAG_MUL_EXT = 8, TAG_POSEIDON_MDS = 10,
              TAG_RANDOM_ACCESS = 13, TAG_REDUCING_EXT = 14, TAG_REDUCING = 15;
__host__ __device__ constexpr bool tag_is_recursion(u32 t) {
    return t == TAG_ARITHMETIC_EXT || t == TAG_COSET_INTERP || t == TAG_EXPONENTIATION || t == TAG_MUL_EXT || t == TAG_POSEIDON_MDS ||
           t == TAG_RANDOM_ACCESS || t == TAG_REDUCING_EXT || t == TAG_REDUCING;
}
// F_{p^2} = F_p[X] / (X^2 - 7) on canonical components: in the prover's base-field evaluation a gate's "extension algebra"
// wires (pairs of columns) are plain F_{p^2} elements
struct QE { u64 a, b; };
ZKB_D QE qe_add(QE x, QE y) { return QE{f_add(x.a, y.a), f_add(x.b, y.b)}; }
ZKB_D QE qe_sub(QE x, QE y) { return QE{f_sub(x.a, y.a), f_sub(x.b, y.b)}; }
ZKB_D QE qe_scale(QE x, u64 s) { return QE{f_mul(x.a, s), f_mul(x.b, s)}; }
__device__ __noinline__ QE qe_mul(QE x, QE y) {
    const u64 bb = f_mul(x.b, y.b);
    return QE{f_add(f_mul(x.a, y.a), f_mul(bb, GL_W)), f_add(f_mul(x.a, y.b), f_mul(x.b, y.a))};
}

struct QuotientArgs {
    const QuotientParams* p;
    const u64* apow;            // [nch][nterms] powers of alpha
    int nterms;
    const u64* cs; size_t cs_stride;
    const u64* w; size_t w_stride;
    const u64* z; size_t z_stride;
    u64* out; size_t out_stride;
    // sliced mode (small circuits, part != nullptr): the launch's work items are spread over blockIdx.y and every slice
    // stores its raw sums in its own slot part[(slot_base + blockIdx.y)][ch][N]; quotient_combine_kernel adds the slots
    u64* part; int slot_base;
    // witness self-check (ZKB_CHECK_WITNESS): the same constraint code evaluated on the SUBGROUP H instead of the LDE coset —
    // cs / w / z are then VALUES over H in natural order (stride n), every filtered constraint sum must be zero, and the last
    // launch raises bit 1 of *unsat_flag instead of dividing by Z_H (which vanishes on H)
    int on_h; unsigned* unsat_flag;
};

// (g0, g1) += c * (p0, p1): deliberately NOT inlined — the quotient kernel has ~150 call sites and its straight-line
// code was 186 KB, far beyond the instruction cache (ncu: stall_no_instruction 0.94, ICC hit rate 85 %).
__device__ __noinline__ ulonglong2 q_acc2(ulonglong2 g, u64 c, u64 p0, u64 p1) {
    g.x = f_add(g.x, f_mul(c, p0));
    g.y = f_add(g.y, f_mul(c, p1));
    return g;
}
__device__ __noinline__ u64 q_sbox7(u64 x) { return gl_sbox7(x); }
// The work of a point is split over launches so that each has a small register and code footprint:
// PART 1 = L0 term + permutation checks + the simple gates (stores the raw sums), PART 3 = the recursion gate set (adds its
// sums; launched only for circuits that have such gates), PART 2 = Poseidon gates (adds its sums and divides by Z_H).
// The constraint-to-alpha-power mapping is the same in all of them.
template <int PART>
__global__ void __launch_bounds__(128, PART == 3 ? 4 : 8) quotient_kernel(QuotientArgs a) {
    const QuotientParams& P = *a.p;
    const unsigned lgN = a.on_h ? P.lg_n : P.lg_n + P.rate_bits;
    const size_t N = size_t(1) << lgN;
    size_t l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (l >= N) return;
    const u32 i = a.on_h ? (u32)l : bitrev32((u32)l, lgN);
    const u32 rate_mask = (1u << P.rate_bits) - 1;
    const size_t l_next = a.on_h ? ((l + 1) & (N - 1)) : bitrev32((u32)((i + (1u << P.rate_bits)) & (N - 1)), lgN);
    const u64 x = a.on_h ? root_pow_lg(lgN, i, false) : f_mul(GL_GEN, root_pow_lg(lgN, i, false));
    const int nch = P.num_challenges, npp = P.num_partial_products, nchunks = npp + 1, chunk = P.qdf;
    const u64* cs = a.cs + l;
    const u64* w = a.w + l;
    const u64* z = a.z + l;
    const u64* apow0 = a.apow;
    const u64* apow1 = a.apow + a.nterms;
    u64 acc0 = 0, acc1 = 0;      // results for challenge 0 / 1 (nch <= 2 supported)
    int term = 0;
    auto add_term = [&](u64 t) {
        ulonglong2 g = q_acc2(make_ulonglong2(acc0, acc1), t, apow0[term], nch > 1 ? apow1[term] : 0);
        acc0 = g.x; acc1 = g.y;
        ++term;
    };
    const int nterm_perm = nch * (1 + nchunks);     // L0 terms + partial-product checks come first
    // sliced mode: PART 1 = one slice per challenge's permutation terms + a last slice for its gates; PART 3 = its gates
    // round-robin over the slices
    const int slice = blockIdx.y, nsl = gridDim.y;
    if (PART == 1) {
    // L0(x) (Z - 1)
    // on H: L0 is the indicator of the first row
    u64 l0 = a.on_h ? (u64)(i == 0) : f_mul(P.zh[i & rate_mask], gl_inv(f_mul(gl_canon(u64(1) << P.lg_n), f_sub(x, 1))));
    for (int ch = 0; ch < nch; ++ch) {
        if (nsl > 1 && slice != ch) continue;
        term = ch;
        add_term(f_mul(l0, f_sub(z[(size_t)ch * a.z_stride], 1)));
    }
    // partial product checks
    if (nsl == 1 && nch == 2) {
        // both challenges in one pass over the routed wires: every wire and sigma value is loaded once and the four
        // running products are independent multiply chains. (Measured: the same 0.33 ms as the per-challenge form below —
        // this loop is 54 % of the launch's instructions, 640 modular multiplies per point, and runs at the ~55 % issue
        // ceiling of IMAD.WIDE + carry-chain code, so the saved loads do not show; profiles/r01_ncu_quotient_v3.md)
        const u64 b0 = P.betas[0], b1 = P.betas[1], gm0 = P.gammas[0], gm1 = P.gammas[1];
        const u64 bx0 = f_mul(b0, x), bx1 = f_mul(b1, x);
        u64 prev0 = z[0], prev1 = z[a.z_stride];
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k) {
            const u64 next0 = k < npp ? z[(size_t)(nch + k) * a.z_stride] : a.z[l_next];
            const u64 next1 = k < npp ? z[(size_t)(nch + npp + k) * a.z_stride] : a.z[a.z_stride + l_next];
            u64 num0 = 1, den0 = 1, num1 = 1, den1 = 1;
            const int j1 = min(P.num_routed, (k + 1) * chunk);
#pragma unroll 2
            for (int j = k * chunk; j < j1; ++j) {
                const u64 wj = w[(size_t)j * a.w_stride], sj = cs[(size_t)(P.num_constants + j) * a.cs_stride], kj = P.k_is[j];
                const u64 wg0 = f_add(wj, gm0), wg1 = f_add(wj, gm1);
                num0 = f_mul(num0, f_add(wg0, f_mul(bx0, kj)));
                den0 = f_mul(den0, f_add(wg0, f_mul(b0, sj)));
                num1 = f_mul(num1, f_add(wg1, f_mul(bx1, kj)));
                den1 = f_mul(den1, f_add(wg1, f_mul(b1, sj)));
            }
            term = nch + k;
            add_term(f_sub(f_mul(prev0, num0), f_mul(next0, den0)));
            term = nch + nchunks + k;
            add_term(f_sub(f_mul(prev1, num1), f_mul(next1, den1)));
            prev0 = next0; prev1 = next1;
        }
    } else {
    for (int ch = 0; ch < nch; ++ch) {
        if (nsl > 1 && slice != ch) continue;
        term = nch + ch * nchunks;
        u64 beta = P.betas[ch], gamma = P.gammas[ch];
        u64 bx = f_mul(beta, x);
        u64 prev = z[(size_t)ch * a.z_stride];
#pragma unroll 1
        for (int k = 0; k < nchunks; ++k) {
            u64 next = k < npp ? z[(size_t)(nch + ch * npp + k) * a.z_stride] : a.z[(size_t)ch * a.z_stride + l_next];
            u64 num = 1, den = 1;
            int j1 = min(P.num_routed, (k + 1) * chunk);
#pragma unroll 1
            for (int j = k * chunk; j < j1; ++j) {
                u64 wg = f_add(w[(size_t)j * a.w_stride], gamma);
                num = f_mul(num, f_add(wg, f_mul(bx, P.k_is[j])));
                den = f_mul(den, f_add(wg, f_mul(beta, cs[(size_t)(P.num_constants + j) * a.cs_stride])));
            }
            add_term(f_sub(f_mul(prev, num), f_mul(next, den)));
            prev = next;
        }
    }
    }
    }
    // gate constraints, filtered
    const int goff = nterm_perm;
    const int nsel = P.num_selectors;
    int ordinal = 0;             // index of the gate among this launch's gates
    for (int g = 0; g < P.num_gates; ++g) {
        const GateDesc gd = P.gates[g];
        if ((gd.tag == TAG_POSEIDON ? 2 : tag_is_recursion(gd.tag) ? 3 : 1) != PART) continue;
        if (nsl > 1 && (PART == 1 ? slice != nsl - 1 : (ordinal++ % nsl) != slice)) continue;
        u64 s = cs[(size_t)gd.selector_index * a.cs_stride];
        u64 filter = 1;
#pragma unroll 1
        for (u32 r = gd.group_lo; r < gd.group_hi; ++r)
            if (r != gd.row) filter = f_mul(filter, f_sub((u64)r, s));
        if (nsel > 1) filter = f_mul(filter, f_sub(0xFFFFFFFFULL, s));
        u64 g0 = 0, g1 = 0;
        int k = goff;
        auto add_c = [&](u64 c) {     // c lazy
            ulonglong2 g = q_acc2(make_ulonglong2(g0, g1), c, apow0[k], nch > 1 ? apow1[k] : 0);
            g0 = g.x; g1 = g.y;
            ++k;
        };
        auto W = [&](int j) { return w[(size_t)j * a.w_stride]; };
        auto K = [&](int j) { return cs[(size_t)(nsel + j) * a.cs_stride]; };
        switch (gd.tag) {
            case TAG_NOOP: break;
            case TAG_CONSTANT:
                if (PART != 1) break;      // PART is a template constant: the other launches' cases compile to nothing
#pragma unroll 1
                for (u32 j = 0; j < gd.param; ++j) add_c(f_sub(K(j), W(j)));
                break;
            case TAG_PUBLIC_INPUT:
                if (PART != 1) break;
                for (int j = 0; j < 4; ++j) add_c(f_sub(W(j), P.pi_hash[j]));
                break;
            case TAG_BASE_SUM: {
                if (PART != 1) break;
                u64 sum = 0;
#pragma unroll 4
                for (int j = (int)gd.param; j-- > 0;) sum = f_add(f_add(sum, sum), W(1 + j));
                add_c(f_sub(sum, W(0)));
#pragma unroll 1
                for (u32 j = 0; j < gd.param; ++j) { u64 b = W(1 + j); add_c(f_mul(b, f_sub(b, 1))); }
                break;
            }
            case TAG_ARITHMETIC: {
                if (PART != 1) break;
                u64 c0 = K(0), c1 = K(1);
#pragma unroll 1
                for (u32 j = 0; j < gd.param; ++j) {
                    u64 prod = f_mul(f_mul(W(4 * j), W(4 * j + 1)), c0);
                    add_c(f_sub(W(4 * j + 3), f_add(prod, f_mul(W(4 * j + 2), c1))));
                }
                break;
            }
            case TAG_POSEIDON: {
                if (PART != 2) break;
                u64 swap = W(24);
                add_c(f_mul(swap, f_sub(swap, 1)));
                u64 st[12];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    u64 lhs = W(j), rhs = W(j + 4), d = W(25 + j);
                    add_c(f_sub(f_mul(swap, f_sub(rhs, lhs)), d));
                    st[j] = f_add(lhs, d);
                    st[j + 4] = f_sub(rhs, d);
                }
#pragma unroll
                for (int j = 8; j < 12; ++j) st[j] = W(j);
                // The state lives in limb form as in the hash kernels (poseidon.cuh): the MDS layer is the shift/add circulant
                // on three 22-bit limb vectors with the NEXT round's constants folded into its output adds, so after it the
                // limbs hold exactly "state + round constants" — the value every constraint of the next round compares with its
                // S-box-input wire. Only the words a constraint needs are folded to a 64-bit value (limb_to_u64_biased); in
                // the 22 partial rounds that is one word, the other 11 are just carry-normalised. (The first version ran a
                // 64-bit-word MDS, 288 IMAD.WIDE per round, and was bound by the FMA pipe.) One loop over the 30 rounds,
                // full rounds in 3 passes of 4 words with a register rotation, non-inlined S-box: small code.
                // wires: full rounds 1..3 -> 29 + 12 (r - 1) + j, partial r' -> 65 + r', last four -> 87 + 12 r'' + j
                u32 o0[12], o1[12], o2[12];
#pragma unroll
                for (int j = 0; j < 12; ++j) {
                    limb_split(f_canon(st[j]), o0[j], o1[j], o2[j]);
                    o0[j] += c_rc3[3 * j]; o1[j] += c_rc3[3 * j + 1]; o2[j] += c_rc3[3 * j + 2];
                }
                auto rotate4 = [&]() {
                    u32 t;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        t = o0[i]; o0[i] = o0[i + 4]; o0[i + 4] = o0[i + 8]; o0[i + 8] = t;
                        t = o1[i]; o1[i] = o1[i + 4]; o1[i + 4] = o1[i + 8]; o1[i + 8] = t;
                        t = o2[i]; o2[i] = o2[i + 4]; o2[i + 4] = o2[i + 8]; o2[i + 8] = t;
                    }
                };
#pragma unroll 1
                for (int r = 0; r < P_ROUNDS; ++r) {
                    if (r < P_HALF_FULL || r >= P_HALF_FULL + P_PARTIAL) {
                        const int base = r < P_HALF_FULL ? 29 + 12 * (r - 1) : 87 + 12 * (r - P_HALF_FULL - P_PARTIAL);
#pragma unroll 1
                        for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                u64 v = limb_to_u64_biased(o0[i], o1[i], o2[i]);
                                if (r != 0) {
                                    const u64 sin = W(base + 4 * pass + i);
                                    add_c(gl_sub_lazy_c(v, sin));
                                    v = sin;
                                }
                                limb_split(q_sbox7(v), o0[i], o1[i], o2[i]);
                            }
                            rotate4();
                        }
                    } else {
                        const u64 sin = W(65 + r - P_HALF_FULL);
                        add_c(gl_sub_lazy_c(limb_to_u64_biased(o0[0], o1[0], o2[0]), sin));
                        limb_split(q_sbox7(sin), o0[0], o1[0], o2[0]);
#pragma unroll
                        for (int j = 1; j < 12; ++j) limb_normalize(o0[j], o1[j], o2[j]);
                    }
                    const u32* rc = c_rc3 + 36 * (r + 1);
                    mds_limb12(o0, rc);
                    mds_limb12(o1, rc + 1);
                    mds_limb12(o2, rc + 2);
                }
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) add_c(gl_sub_lazy_c(limb_to_u64_biased(o0[i], o1[i], o2[i]), W(12 + 4 * pass + i)));
                    rotate4();
                }
                break;
            }
            case TAG_ARITHMETIC_EXT:       // per op: a[2] b[2] addend[2] out[2];  out - (c0 a b + c1 addend)
            case TAG_MUL_EXT: {            // per op: a[2] b[2] out[2];            out - c0 a b
                if (PART != 3) break;
                const bool addend = gd.tag == TAG_ARITHMETIC_EXT;
                const int per = addend ? 8 : 6;
                const u64 c0 = K(0), c1 = addend ? K(1) : 0;
#pragma unroll 1
                for (u32 j = 0; j < gd.param; ++j) {
                    const int b = per * (int)j;
                    QE t = qe_scale(qe_mul(QE{W(b), W(b + 1)}, QE{W(b + 2), W(b + 3)}), c0);
                    if (addend) t = qe_add(t, qe_scale(QE{W(b + 4), W(b + 5)}, c1));
                    const QE c = qe_sub(QE{W(b + per - 2), W(b + per - 1)}, t);
                    add_c(c.a); add_c(c.b);
                }
                break;
            }
            case TAG_REDUCING:             // out[0..2) alpha[2..4) old_acc[4..6) coeffs, then the intermediate accumulators;
            case TAG_REDUCING_EXT: {       // acc_i = acc_{i-1} alpha + coeff_i, the last accumulator is the output
                if (PART != 3) break;
                const bool ext = gd.tag == TAG_REDUCING_EXT;
                const int nc = (int)gd.param, start_accs = 6 + (ext ? 2 * nc : nc);
                const QE alpha{W(2), W(3)};
                QE acc{W(4), W(5)};
#pragma unroll 1
                for (int j = 0; j < nc; ++j) {
                    QE t = qe_mul(acc, alpha);
                    if (ext) t = qe_add(t, QE{W(6 + 2 * j), W(7 + 2 * j)});
                    else t.a = f_add(t.a, W(6 + j));
                    const int at = j + 1 == nc ? 0 : start_accs + 2 * j;
                    acc = QE{W(at), W(at + 1)};
                    const QE c = qe_sub(t, acc);          // upstream: acc * alpha + coeff - accs[i]
                    add_c(c.a); add_c(c.b);
                }
                break;
            }
            case TAG_RANDOM_ACCESS: {      // per copy: index, claimed element, 2^bits items; extra constants; then the bit wires
                if (PART != 3) break;
                const int bits = (int)gd.param, vec = 1 << bits, copies = (int)gd.p2, extra = (int)gd.p3;
                const int routed = (2 + vec) * copies + extra;
#pragma unroll 1
                for (int cpy = 0; cpy < copies; ++cpy) {
                    const int base = (2 + vec) * cpy, wb = routed + cpy * bits;
                    u64 idx = 0;
                    for (int j = 0; j < bits; ++j) { const u64 b = W(wb + j); add_c(f_mul(b, f_sub(b, 1))); }
                    for (int j = bits; j-- > 0;) idx = f_add(f_add(idx, idx), W(wb + j));
                    add_c(f_sub(idx, W(base)));
                    // fold the list pairwise, level j with bit j: x + b (y - x); streamed through a stack of one partial
                    // result per level so that the items are read once and never all live at once
                    u64 stack[6];
#pragma unroll 1
                    for (int k = 0; k < vec; k += 2) {
                        u64 x = W(base + 2 + k), y = W(base + 3 + k);
                        u64 v = f_add(x, f_mul(W(wb), f_sub(y, x)));
                        int lvl = 1;
                        for (int m = k >> 1; m & 1; m >>= 1, ++lvl) v = f_add(stack[lvl], f_mul(W(wb + lvl), f_sub(v, stack[lvl])));
                        stack[lvl] = v;
                    }
                    add_c(f_sub(stack[bits], W(base + 1)));
                }
                for (int j = 0; j < extra; ++j) add_c(f_sub(K(j), W((2 + vec) * copies + j)));
                break;
            }
            case TAG_EXPONENTIATION: {     // base 0, power bits 1..nb (little endian), output nb+1, intermediates nb+2..
                if (PART != 3) break;
                const int nb = (int)gd.param;
                const u64 bm1 = f_sub(W(0), 1);
                u64 prev = 1;
#pragma unroll 1
                for (int j = 0; j < nb; ++j) {
                    const u64 factor = f_add(f_mul(W(nb - j), bm1), 1);      // bit * base + (1 - bit)
                    const u64 cur = W(nb + 2 + j);
                    add_c(f_sub(f_mul(prev, factor), cur));
                    prev = f_mul(cur, cur);
                }
                add_c(f_sub(W(nb + 1), W(2 * nb + 1)));
                break;
            }
            case TAG_COSET_INTERP: {       // shift 0; 2^bits values; point; value; intermediates (evals, then products); shifted point
                if (PART != 3) break;
                const int np = 1 << gd.param, deg = (int)gd.p2, ni = (np - 2) / (deg - 1);
                const int at_point = 1 + 2 * np, at_value = at_point + 2, at_inter = at_value + 2, at_shifted = at_inter + 4 * ni;
                const QE shifted{W(at_shifted), W(at_shifted + 1)};
                {
                    const QE c = qe_sub(QE{W(at_point), W(at_point + 1)}, qe_scale(shifted, W(0)));
                    add_c(c.a); add_c(c.b);
                }
                QE eval{0, 0}, prod{1, 0};
                int lo = 0, hi = deg;
#pragma unroll 1
                for (int seg = 0; seg <= ni; ++seg) {
#pragma unroll 1
                    for (int k = lo; k < hi; ++k) {      // eval <- eval (z - x_k) + w_k v_k prod;  prod <- prod (z - x_k)
                        const QE term{f_sub(shifted.a, P.bary_x[k]), shifted.b};
                        const QE wv = qe_scale(QE{W(1 + 2 * k), W(2 + 2 * k)}, P.bary_w[k]);
                        eval = qe_add(qe_mul(eval, term), qe_mul(wv, prod));
                        prod = qe_mul(prod, term);
                    }
                    if (seg == ni) break;
                    const QE ie{W(at_inter + 2 * seg), W(at_inter + 2 * seg + 1)};
                    const QE ip{W(at_inter + 2 * (ni + seg)), W(at_inter + 2 * (ni + seg) + 1)};
                    QE c = qe_sub(ie, eval);
                    add_c(c.a); add_c(c.b);
                    c = qe_sub(ip, prod);
                    add_c(c.a); add_c(c.b);
                    eval = ie; prod = ip;
                    lo = 1 + (deg - 1) * (seg + 1);
                    hi = min(lo + deg - 1, np);
                }
                const QE c = qe_sub(QE{W(at_value), W(at_value + 1)}, eval);
                add_c(c.a); add_c(c.b);
                break;
            }
            case TAG_POSEIDON_MDS: {       // 12 F_{p^2} inputs at 2 i, outputs at 24 + 2 i: the MDS matrix acts on each component
                if (PART != 3) break;
                // one component at a time (12 live words instead of 24); constraint 2 j + comp gets alpha^(k + 2 j + comp)
#pragma unroll 1
                for (int comp = 0; comp < 2; ++comp) {
                    u64 sv[12];
#pragma unroll
                    for (int j = 0; j < 12; ++j) sv[j] = W(2 * j + comp);
                    mds_layer(sv);
#pragma unroll
                    for (int j = 0; j < 12; ++j) {
                        const u64 c = f_sub(W(24 + 2 * j + comp), f_canon(sv[j]));
                        const ulonglong2 g = q_acc2(make_ulonglong2(g0, g1), c, apow0[k + 2 * j + comp], nch > 1 ? apow1[k + 2 * j + comp] : 0);
                        g0 = g.x; g1 = g.y;
                    }
                }
                k += 24;
                break;
            }
            default: break;   // rejected on the host (ZKB_E_UNSUPPORTED_GATE)
        }
        acc0 = f_add(acc0, f_mul(filter, g0));
        if (nch > 1) acc1 = f_add(acc1, f_mul(filter, g1));
    }
    if (a.part) {
        u64* dst = a.part + ((size_t)(a.slot_base + slice) * nch << lgN) + l;
        dst[0] = acc0;
        if (nch > 1) dst[N] = acc1;
    } else if (PART == 1) {
        a.out[l] = acc0;
        if (nch > 1) a.out[a.out_stride + l] = acc1;
    } else if (PART == 3) {
        a.out[l] = f_add(acc0, a.out[l]);
        if (nch > 1) a.out[a.out_stride + l] = f_add(acc1, a.out[a.out_stride + l]);
    } else if (a.on_h) {
        const u64 v0 = f_add(acc0, a.out[l]), v1 = nch > 1 ? f_add(acc1, a.out[a.out_stride + l]) : 0;
        if (v0 | v1) atomicOr(a.unsat_flag, 2u);
    } else {
        u64 zi = P.zh_inv[i & rate_mask];
        a.out[l] = f_mul(f_add(acc0, a.out[l]), zi);
        if (nch > 1) a.out[a.out_stride + l] = f_mul(f_add(acc1, a.out[a.out_stride + l]), zi);
    }
}

// out[ch][l] = (sum over slots of part[slot][ch][l]) / Z_H
__global__ void quotient_combine_kernel(const QuotientParams* p, const u64* part, int nslots, u64* out, size_t out_stride) {
    const QuotientParams& P = *p;
    const unsigned lgN = P.lg_n + P.rate_bits;
    const size_t N = size_t(1) << lgN, l = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (l >= N) return;
    const int ch = blockIdx.y, nch = P.num_challenges;
    u64 acc = 0;
    for (int s = 0; s < nslots; ++s) acc = f_add(acc, part[((size_t)(s * nch + ch) << lgN) + l]);
    const u32 i = bitrev32((u32)l, lgN);
    out[(size_t)ch * out_stride + l] = f_mul(acc, P.zh_inv[i & ((1u << P.rate_bits) - 1)]);
}

int quotient_slots(const QuotientParams& ph) {
    int rec = 0;
    for (int g = 0; g < ph.num_gates; ++g) rec += tag_is_recursion(ph.gates[g].tag);
    return (ph.num_challenges + 1) + rec + 1;
}

void launch_quotient(const QuotientParams* params_dev, const QuotientParams& ph, const u64* apow_dev, int nterms,
                     const u64* cs_lde, size_t cs_stride, const u64* wires_lde, size_t w_stride, const u64* zs_lde,
                     size_t z_stride, u64* out, size_t out_stride, cudaStream_t st, const QuotientFork* fork) {
    QuotientArgs a{params_dev, apow_dev, nterms, cs_lde, cs_stride, wires_lde, w_stride, zs_lde, z_stride, out, out_stride, nullptr, 0,
                   0, nullptr};
    size_t N = size_t(1) << (ph.lg_n + ph.rate_bits);
    const unsigned gx = (unsigned)((N + 127) / 128);
    if (fork && fork->part) {
        // small circuit: one thread per point leaves most of the GPU idle (2^15 points = 7 warps per SM) and a thread's gate
        // list is one long dependency chain, so the three launches run CONCURRENTLY on the circuit's stream and two helper
        // streams, their work items spread over blockIdx.y, each slice writing its own slot; one small launch adds the slots
        a.part = fork->part;
        int rec = 0;
        for (int g = 0; g < ph.num_gates; ++g) rec += tag_is_recursion(ph.gates[g].tag);
        const int s1 = ph.num_challenges + 1;
        ZKB_CUDA_CHECK(cudaEventRecord(fork->fork, st));
        for (int k = 0; k < 2; ++k) ZKB_CUDA_CHECK(cudaStreamWaitEvent(fork->aux[k], fork->fork, 0));
        // launch order = CTA dispatch order: the Poseidon launch (one long chain per thread, 256 CTAs at 2^15 points) and
        // part 1 become resident together; the recursion slices (most CTAs, 128 registers) fill in behind them
        a.slot_base = s1 + rec;
        ZKB_COUNT_LAUNCH();
        quotient_kernel<2><<<dim3(gx, 1), 128, 0, st>>>(a);
        a.slot_base = 0;
        ZKB_COUNT_LAUNCH();
        quotient_kernel<1><<<dim3(gx, s1), 128, 0, fork->aux[0]>>>(a);
        if (rec) {
            a.slot_base = s1;
            ZKB_COUNT_LAUNCH();
            quotient_kernel<3><<<dim3(gx, rec), 128, 0, fork->aux[1]>>>(a);
        }
        for (int k = 0; k < 2; ++k) {
            ZKB_CUDA_CHECK(cudaEventRecord(fork->join[k], fork->aux[k]));
            ZKB_CUDA_CHECK(cudaStreamWaitEvent(st, fork->join[k], 0));
        }
        ZKB_COUNT_LAUNCH();
        quotient_combine_kernel<<<dim3(gx, ph.num_challenges), 128, 0, st>>>(params_dev, fork->part, s1 + rec + 1, out, out_stride);
        return;
    }
    ZKB_COUNT_LAUNCH();
    quotient_kernel<1><<<gx, 128, 0, st>>>(a);
    if (ph.has_recursion_gates) {
        ZKB_COUNT_LAUNCH();
        quotient_kernel<3><<<gx, 128, 0, st>>>(a);
    }
    ZKB_COUNT_LAUNCH();
    quotient_kernel<2><<<gx, 128, 0, st>>>(a);
}

// the witness self-check: all three launches over the n points of H (values, natural order), flag bit 1 on a violation
void launch_constraint_check(const QuotientParams* params_dev, const QuotientParams& ph, const u64* apow_dev, int nterms,
                             const u64* cs_vals, size_t cs_stride, const u64* wires_vals, size_t w_stride, const u64* zs_vals,
                             size_t z_stride, u64* scratch, size_t scratch_stride, unsigned* flag_dev, cudaStream_t st) {
    QuotientArgs a{params_dev, apow_dev, nterms, cs_vals, cs_stride, wires_vals, w_stride, zs_vals, z_stride, scratch, scratch_stride,
                   nullptr, 0, 1, flag_dev};
    const size_t n = size_t(1) << ph.lg_n;
    const unsigned gx = (unsigned)((n + 127) / 128);
    ZKB_COUNT_LAUNCH();
    quotient_kernel<1><<<gx, 128, 0, st>>>(a);
    if (ph.has_recursion_gates) {
        ZKB_COUNT_LAUNCH();
        quotient_kernel<3><<<gx, 128, 0, st>>>(a);
    }
    ZKB_COUNT_LAUNCH();
    quotient_kernel<2><<<gx, 128, 0, st>>>(a);
}

// ---------------------------------------------------------------------------------------------
// Openings (a9): P(z) = sum_k c_k z^k for many polynomials at one extension point
// ---------------------------------------------------------------------------------------------
__global__ void ext_powers_kernel(ext2 z, unsigned lg_n, u64* pa, u64* pb) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= (size_t(1) << lg_n)) return;
    ext2 r = e_pow(z, k);
    pa[k] = r.a;
    pb[k] = r.b;
}
void launch_ext_powers(ext2 z, unsigned lg_n, u64* pa, u64* pb, cudaStream_t st) {
    size_t n = size_t(1) << lg_n;
    ZKB_COUNT_LAUNCH();
    ext_powers_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(z, lg_n, pa, pb);
}

__global__ void __launch_bounds__(256) eval_polys_kernel(const u64* __restrict__ coeffs, size_t stride, unsigned lg_n,
                                                         const u64* __restrict__ za, const u64* __restrict__ zb, u64* __restrict__ out) {
    __shared__ u64 sa[8], sb[8];
    const u64* c = coeffs + (size_t)blockIdx.x * stride;
    size_t n = size_t(1) << lg_n;
    u64 a = 0, b = 0;
    for (size_t k = threadIdx.x; k < n; k += blockDim.x) {
        u64 v = c[k];
        a = gl_add(a, gl_mul(v, za[k]));
        b = gl_add(b, gl_mul(v, zb[k]));
    }
    for (int off = 16; off > 0; off >>= 1) {
        a = gl_add(a, __shfl_down_sync(0xffffffffu, a, off));
        b = gl_add(b, __shfl_down_sync(0xffffffffu, b, off));
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sa[warp] = a; sb[warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wv = 1; wv < (int)(blockDim.x >> 5); ++wv) { a = gl_add(a, sa[wv]); b = gl_add(b, sb[wv]); }
        out[2 * blockIdx.x] = a;
        out[2 * blockIdx.x + 1] = b;
    }
}
void launch_eval_polys(const u64* coeffs, size_t stride, int ncols, unsigned lg_n, const u64* za, const u64* zb, u64* out, cudaStream_t st) {
    if (ncols <= 0) return;
    ZKB_COUNT_LAUNCH();
    eval_polys_kernel<<<ncols, 256, 0, st>>>(coeffs, stride, lg_n, za, zb, out);
}

// all openings of a proof in two launches: powers of both points, then one CTA per (segment, polynomial)
__global__ void ext_powers2_kernel(ext2 z0, ext2 z1, unsigned lg_n, u64* pw) {     // pw: [z0.a | z0.b | z1.a | z1.b], n each
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x, n = size_t(1) << lg_n;
    if (k >= n) return;
    ext2 r = e_pow(blockIdx.y ? z1 : z0, k);
    pw[(2 * blockIdx.y) * n + k] = r.a;
    pw[(2 * blockIdx.y + 1) * n + k] = r.b;
}
__global__ void __launch_bounds__(256) eval_polys_multi_kernel(OpeningsArgs a) {
    __shared__ u64 sa[8], sb[8];
    int c = blockIdx.x, seg = 0;
    while (c >= a.seg[seg].ncols) { c -= a.seg[seg].ncols; ++seg; }
    const size_t n = size_t(1) << a.lg_n;
    const u64* cf = a.seg[seg].coeffs + (size_t)c * a.seg[seg].stride;
    const u64* za = a.pw + (size_t)(2 * a.seg[seg].point) * n;
    const u64* zb = za + n;
    u64 x = 0, y = 0;
    for (size_t k = threadIdx.x; k < n; k += blockDim.x) {
        u64 v = cf[k];
        x = gl_add(x, gl_mul(v, za[k]));
        y = gl_add(y, gl_mul(v, zb[k]));
    }
    for (int off = 16; off > 0; off >>= 1) {
        x = gl_add(x, __shfl_down_sync(0xffffffffu, x, off));
        y = gl_add(y, __shfl_down_sync(0xffffffffu, y, off));
    }
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sa[warp] = x; sb[warp] = y; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wv = 1; wv < (int)(blockDim.x >> 5); ++wv) { x = gl_add(x, sa[wv]); y = gl_add(y, sb[wv]); }
        a.out[2 * blockIdx.x] = x;
        a.out[2 * blockIdx.x + 1] = y;
    }
}
void launch_openings(const OpeningsArgs& a, ext2 z0, ext2 z1, cudaStream_t st) {
    const size_t n = size_t(1) << a.lg_n;
    int total = 0;
    for (int s = 0; s < a.nseg; ++s) total += a.seg[s].ncols;
    dim3 g((unsigned)((n + 127) / 128), 2);
    ZKB_COUNT_LAUNCH();
    ext_powers2_kernel<<<g, 128, 0, st>>>(z0, z1, a.lg_n, a.pw);
    if (total <= 0) return;
    ZKB_COUNT_LAUNCH();
    eval_polys_multi_kernel<<<total, 256, 0, st>>>(a);
}

// ---------------------------------------------------------------------------------------------
// FRI (a10-a13)
// ---------------------------------------------------------------------------------------------
struct FriCombineArgs {
    FriCombineParams p;
    const u64* apa; const u64* apb;     // alpha^j, j < total columns
    ext2 alpha_shift;                   // alpha^(#polys opened at g*zeta)
    u64* oa; u64* ob;
};
// 32 points x 8 column groups per CTA: each thread sums every 8th column of its point, the groups are reduced through
// shared memory, and one thread per point finishes with ONE field inversion for both quotients (a/(x - zeta) and
// b/(x - g zeta) share inv(norm_0 norm_1)).
__global__ void __launch_bounds__(256) fri_combine_kernel(FriCombineArgs a) {
    __shared__ u64 red[4][8][32];
    const FriCombineParams& P = a.p;
    const size_t n = size_t(1) << P.lg_n;
    const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t l = blockIdx.x * (size_t)32 + tx;
    u64 s0a = 0, s0b = 0, s1a = 0, s1b = 0;
    if (l < n) {
        int j0 = 0;
        for (int t = 0; t < 4; ++t) {
            const u64* col = P.lde[t] + l;
            const int nc = P.ncols[t];
            int c = (int)((ty + 8 - (j0 & 7)) & 7);              // first column of this batch with (j0 + c) % 8 == ty
            for (; c < nc; c += 8) {
                const u64 v = col[(size_t)c * P.stride[t]];
                const int j = j0 + c;
                s0a = gl_add(s0a, gl_mul(v, a.apa[j]));
                s0b = gl_add(s0b, gl_mul(v, a.apb[j]));
            }
            j0 += nc;
        }
        for (int c = (int)ty; c < P.num_zs; c += 8) {
            const u64 v = P.lde[2][(size_t)c * P.stride[2] + l];
            s1a = gl_add(s1a, gl_mul(v, a.apa[c]));
            s1b = gl_add(s1b, gl_mul(v, a.apb[c]));
        }
    }
    red[0][ty][tx] = s0a; red[1][ty][tx] = s0b; red[2][ty][tx] = s1a; red[3][ty][tx] = s1b;
    __syncthreads();
    if (ty != 0 || l >= n) return;
#pragma unroll
    for (int g = 1; g < 8; ++g) {
        s0a = gl_add(s0a, red[0][g][tx]); s0b = gl_add(s0b, red[1][g][tx]);
        s1a = gl_add(s1a, red[2][g][tx]); s1b = gl_add(s1b, red[3][g][tx]);
    }
    const u64 x = gl_mul(GL_GEN, root_pow_lg(P.lg_n, bitrev32((u32)l, P.lg_n), false));
    // 1 / (x - z) = conj(x - z) / norm(x - z), norm(u + v X) = u^2 - 7 v^2
    const ext2 d0 = e_sub(e_from(x), P.zeta), d1 = e_sub(e_from(x), P.zeta_next);
    const u64 n0 = gl_sub(gl_sqr(d0.a), gl_mul(GL_W, gl_sqr(d0.b))), n1 = gl_sub(gl_sqr(d1.a), gl_mul(GL_W, gl_sqr(d1.b)));
    const u64 inv01 = gl_inv(gl_mul(n0, n1));
    const u64 i0 = gl_mul(inv01, n1), i1 = gl_mul(inv01, n0);
    const ext2 inv0 = e_make(gl_mul(d0.a, i0), gl_mul(gl_neg(d0.b), i0));
    const ext2 inv1 = e_make(gl_mul(d1.a, i1), gl_mul(gl_neg(d1.b), i1));
    ext2 q0 = e_mul(e_sub(e_make(s0a, s0b), P.reduced0), inv0);
    ext2 q1 = e_mul(e_sub(e_make(s1a, s1b), P.reduced1), inv1);
    ext2 q = e_add(e_mul(q0, a.alpha_shift), q1);
    a.oa[l] = q.a;
    a.ob[l] = q.b;
}
void launch_fri_combine(const FriCombineParams& p, const u64* apa, const u64* apb, u64* oa, u64* ob, cudaStream_t st) {
    FriCombineArgs a{p, apa, apb, e_pow(p.alpha, (u64)p.num_zs), oa, ob};
    size_t n = size_t(1) << p.lg_n;
    ZKB_COUNT_LAUNCH();
    fri_combine_kernel<<<(unsigned)((n + 31) / 32), 256, 0, st>>>(a);
}

__global__ void fri_fold_kernel(const u64* __restrict__ ca, const u64* __restrict__ cb, u64* __restrict__ oa, u64* __restrict__ ob,
                                size_t m_out, int arity, ext2 beta) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= m_out) return;
    ext2 acc = e_make(0, 0);
    for (int i = arity; i-- > 0;) acc = e_add(e_mul(acc, beta), e_make(ca[k * arity + i], cb[k * arity + i]));
    oa[k] = acc.a;
    ob[k] = acc.b;
}
void launch_fri_fold(const u64* ca, const u64* cb, u64* oa, u64* ob, size_t m_out, int arity, ext2 beta, cudaStream_t st) {
    if (!m_out) return;
    ZKB_COUNT_LAUNCH();
    fri_fold_kernel<<<(unsigned)((m_out + 127) / 128), 128, 0, st>>>(ca, cb, oa, ob, m_out, arity, beta);
}

// Proof-of-work grind, MIN rule, in ONE launch: thread t tries base + t, base + t + T, ... (T = grid size) and stops
// as soon as its next candidate is not below the smallest witness found so far, so the final *result is the true
// minimum over [base, base + count) no matter how the blocks are scheduled.
__global__ void __launch_bounds__(128) pow_search_kernel(const u64* __restrict__ state12, int pos, u64 base, u64 count, unsigned bits,
                                                         unsigned long long* result) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u64 init[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) init[i] = state12[i];
    for (u64 k = blockIdx.x * (u64)blockDim.x + threadIdx.x; k < count; k += stride) {
        const u64 cand = base + k;
        if (cand >= *reinterpret_cast<volatile unsigned long long*>(result)) break;
        PoseidonState s;
#pragma unroll
        for (int i = 0; i < 12; ++i) s.set(i, i == pos ? cand : init[i]);
        s.permute();
        if (__clzll((long long)s.get(7)) >= (int)bits) atomicMin(result, (unsigned long long)cand);
    }
}
void launch_pow_search(const u64* state12_dev, int pos, u64 base, u64 count, unsigned bits, unsigned long long* result, cudaStream_t st) {
    ZKB_COUNT_LAUNCH();
    u64 blocks = (count + 127) / 128;
    if (blocks > 148 * 4) blocks = 148 * 4;      // ~76 k candidates per sweep: the expected witness (2^16 for 16 bits) falls in the first or second
    pow_search_kernel<<<(unsigned)blocks, 128, 0, st>>>(state12_dev, pos, base, count, bits, result);
}

__global__ void gather_rows_kernel(const u64* __restrict__ lde, size_t stride, int width, const u32* __restrict__ idx, u64* __restrict__ out) {
    int q = blockIdx.x;
    for (int c = threadIdx.x; c < width; c += blockDim.x) out[(size_t)q * width + c] = lde[(size_t)c * stride + idx[q]];
}
void launch_gather_rows(const u64* lde, size_t stride, int width, const u32* idx_dev, int nq, u64* out, cudaStream_t st) {
    if (nq <= 0 || width <= 0) return;
    ZKB_COUNT_LAUNCH();
    gather_rows_kernel<<<nq, 128, 0, st>>>(lde, stride, width, idx_dev, out);
}
__global__ void gather_paths_kernel(const u64* __restrict__ digests, size_t num_leaves, int path_len, const u32* __restrict__ idx,
                                    u64* __restrict__ out) {
    int q = blockIdx.x;
    for (int t = threadIdx.x; t < path_len * 4; t += blockDim.x) {
        int k = t >> 2, e = t & 3;
        size_t off = 0;
        for (int j = 0; j < k; ++j) off += num_leaves >> j;
        size_t node = ((size_t)idx[q] >> k) ^ 1;
        out[((size_t)q * path_len + k) * 4 + e] = digests[(off + node) * 4 + e];
    }
}
void launch_gather_paths(const u64* digests, size_t num_leaves, int path_len, const u32* idx_dev, int nq, u64* out, cudaStream_t st) {
    if (nq <= 0 || path_len <= 0) return;
    ZKB_COUNT_LAUNCH();
    gather_paths_kernel<<<nq, 64, 0, st>>>(digests, num_leaves, path_len, idx_dev, out);
}
__global__ void gather_ext_leaves_kernel(const u64* __restrict__ a, const u64* __restrict__ b, int arity, const u32* __restrict__ idx,
                                         u64* __restrict__ out) {
    int q = blockIdx.x;
    for (int k = threadIdx.x; k < arity; k += blockDim.x) {
        size_t src = (size_t)idx[q] * arity + k;
        out[((size_t)q * arity + k) * 2] = a[src];
        out[((size_t)q * arity + k) * 2 + 1] = b[src];
    }
}
void launch_gather_ext_leaves(const u64* a, const u64* b, int arity, const u32* idx_dev, int nq, u64* out, cudaStream_t st) {
    if (nq <= 0) return;
    ZKB_COUNT_LAUNCH();
    gather_ext_leaves_kernel<<<nq, 32, 0, st>>>(a, b, arity, idx_dev, out);
}

}  // namespace zkb
