// Shared-memory NTT for transform sizes up to 2^14 (one column = one CTA), the size class of the wormhole / voting
// circuits: replaces qp-plonky2-field's `ifft` / `coset_fft` / `lde` on the device.
//
//  * radix-8 / radix-16 decimation-in-frequency passes: a thread loads 8 / 16 elements (stride 2^(s-K)) into registers, runs
//    the butterfly stages there — the internal twiddles w_16^k are powers of two (w_16 = 2^12), applied as shift forms
//    (field.cuh f_shl), no 64x64 multiply — and multiplies by the pass twiddle
//    w_{2^s}^(b * bitrev4(p)) from one table of w_{2^14}^t — 4 block-wide barriers and 4 shared-memory round trips for
//    14 stages instead of 14;
//  * natural order in, bit-reversed order out, which IS the Merkle leaf order of PolynomialBatch, so the LDE is written
//    with coalesced stores and no transpose; the inverse transform uses the same forward code and reads the result at
//    bitrev((n - k) mod n);
//  * the coset pre-scale (shift * w_N^j)^k is one multiply per element at load time from a cached table;
//  * field add/sub on canonical values with carry flags (5 / 7 instructions; the compiler's compare+select form is 9+).
#pragma once
#include "field.cuh"
#include <utility>

namespace zkb {

constexpr unsigned NTT_SM_LG = 14;                       // largest transform held in shared memory
__device__ u64 d_W14[1u << NTT_SM_LG];                   // w_{2^14}^t, 0 <= t < 2^14

ZKB_D unsigned ntt_pad(unsigned i) { return i + (i >> 4); }          // one spare word per 16: keeps small-stride passes off one bank
inline size_t ntt_smem_bytes(unsigned lg) { return sizeof(u64) * ((size_t(1) << lg) + (size_t(1) << lg) / 16 + 1); }

// in-register DIF of 2^K elements (K <= 6); r[p] ends up holding output index bitrev_K(p). The internal twiddles w_64^i = 2^(3 i)
// are powers of two: shift forms (field.cuh f_shl), no 64x64 multiply. INV: inverse twiddles.
template <int K, bool INV, int T, int G, int J>
struct RadixStep {      // butterfly (g + j, g + j + half) of stage T, half = 2^(K-1-T); recursion over j, g, t at compile time
    static ZKB_D void run(u64* r) {
        constexpr int half = 1 << (K - 1 - T);
        const u64 a = r[G + J], b = r[G + J + half];
        r[G + J] = f_add(a, b);
        if constexpr (J == 0) r[G + J + half] = f_sub(a, b);
        else r[G + J + half] = f_sub_twiddle64<J * (32 / half), INV>(a, b);
        if constexpr (J + 1 < half) RadixStep<K, INV, T, G, J + 1>::run(r);
        else if constexpr (G + 2 * half < (1 << K)) RadixStep<K, INV, T, G + 2 * half, 0>::run(r);
        else if constexpr (T + 1 < K) RadixStep<K, INV, T + 1, 0, 0>::run(r);
    }
};
template <int K>
ZKB_D void radix_dif(u64* r, bool inv) {
    if (inv) RadixStep<K, true, 0, 0, 0>::run(r);
    else RadixStep<K, false, 0, 0, 0>::run(r);
}

// One radix-2^K pass over sm[0 .. 2^L). The array is 2^lgC interleaved transforms (element index = row * 2^lgC + c;
// lgC = 0 for a single transform): sub-transforms of 2^(s - lgC) rows, tile = elements b + e * 2^(s-K).
// The pass twiddles w^(row(b) * bitrev(p)) are fetched BEFORE the butterflies so their latency hides behind the arithmetic.
template <int K>
ZKB_D void ntt_dif_pass(u64* sm, unsigned L, unsigned s, unsigned lgC, bool inv) {
    constexpr int R = 1 << K;
    const unsigned lgM = s - K, M = 1u << lgM, ntiles = 1u << (L - K);
    const unsigned tmask = (1u << NTT_SM_LG) - 1;
    for (unsigned t = threadIdx.x; t < ntiles; t += blockDim.x) {
        const unsigned b = t & (M - 1), base = ((t >> lgM) << s) + b;
        u64 tw[R];
        if (lgM > lgC) {
            const unsigned brow = b >> lgC;
#pragma unroll
            for (int p = 1; p < R; ++p) {
                const unsigned q = __brev((unsigned)p) >> (32 - K);
                unsigned idx = (brow * q) << (NTT_SM_LG - (s - lgC));
                if (inv) idx = (0u - idx) & tmask;
                tw[p] = __ldg(&d_W14[idx]);
            }
        }
        u64 r[R];
#pragma unroll
        for (int e = 0; e < R; ++e) r[e] = sm[ntt_pad(base + ((unsigned)e << lgM))];
        radix_dif<K>(r, inv);
        if (lgM > lgC) {
#pragma unroll
            for (int p = 1; p < R; ++p) r[p] = f_mul(r[p], tw[p]);
        }
#pragma unroll
        for (int p = 0; p < R; ++p) sm[ntt_pad(base + ((unsigned)p << lgM))] = r[p];
    }
}
// transform in place (forward, or inverse without the 1/n): natural order in, bit-reversed rows out
// (sm[pad(i * 2^lgC + c)] = X_c[bitrev(i)]). Ends with a barrier. KMAX = 4: radix-16 passes (16 elements + 15 twiddles in
// registers: 128 registers, 512 threads); KMAX = 3: radix-8 passes (one more shared-memory round trip, the same number of
// multiplies, 64 registers so that 1024 threads = 32 warps share the column).
template <int KMAX = 4>
ZKB_D void ntt_dif_smem(u64* sm, unsigned L, unsigned lgC = 0, bool inv = false) {
    unsigned s = L;
    while (s >= lgC + KMAX) { ntt_dif_pass<KMAX>(sm, L, s, lgC, inv); s -= KMAX; __syncthreads(); }
    const unsigned rem = s - lgC;
    if (KMAX > 3 && rem == 3) ntt_dif_pass<3>(sm, L, s, lgC, inv);
    else if (rem == 2) ntt_dif_pass<2>(sm, L, s, lgC, inv);
    else if (rem == 1) ntt_dif_pass<1>(sm, L, s, lgC, inv);
    if (rem) __syncthreads();
}

// ---- the same passes with EVERYTHING known at compile time (transform size L, pass position S, block size): all shared-memory
// offsets become immediates (pad(base + e M) = pad(base) + e M + (e M >> 4): the low four bits never carry, see the table in
// DESIGN.md §4.2), the twiddle-index arithmetic folds, the tile loops unroll, and the passes with stride < 16 remap lanes to
// tiles so that a half-warp's 16 eight-byte accesses fall into 16 different bank pairs (the generic pass had 2-way conflicts
// at stride 4: 15 % of the kernel's shared-memory wavefronts). Forward transforms of the sizes the circuits use (2^12 - 2^14).
template <int K, int S, int L, int THREADS>
ZKB_D void ntt_pass_c(u64* sm) {
    constexpr int R = 1 << K, lgM = S - K, M = 1 << lgM, NT = 1 << (L - K);
    constexpr unsigned tmask = (1u << NTT_SM_LG) - 1;
#pragma unroll
    for (int it = 0; it < (NT + THREADS - 1) / THREADS; ++it) {
        unsigned t = threadIdx.x + it * THREADS;
        if ((NT % THREADS) != 0 && t >= (unsigned)NT) break;
        if (M == 4) {            // half-warp: 8 groups x 2 residues -> bank pairs 2 g + b all different
            const unsigned l = t & 15, half = (t >> 4) & 1;
            t = (t & ~31u) + ((l >> 1) << 2) + (l & 1) + 2 * half;
        } else if (M == 2) {     // half-warp: 16 groups, one residue
            const unsigned l = t & 15, half = (t >> 4) & 1;
            t = (t & ~31u) + (l << 1) + half;
        }
        const unsigned b = t & (M - 1), base = ((t >> lgM) << S) + b, pbase = base + (base >> 4);
        u64 tw[R];
        if (lgM > 0) {
#pragma unroll
            for (int p = 1; p < R; ++p) {
                constexpr int dummy = 0; (void)dummy;
                const unsigned q = __brev((unsigned)p) >> (32 - K);
                tw[p] = __ldg(&d_W14[((b * q) << (NTT_SM_LG - S)) & tmask]);
            }
        }
        u64 r[R];
#pragma unroll
        for (int e = 0; e < R; ++e) r[e] = sm[pbase + e * M + ((e * M) >> 4)];
        radix_dif<K>(r, false);
        if (lgM > 0) {
#pragma unroll
            for (int p = 1; p < R; ++p) r[p] = f_mul(r[p], tw[p]);
        }
#pragma unroll
        for (int p = 0; p < R; ++p) sm[pbase + p * M + ((p * M) >> 4)] = r[p];
    }
}
template <int L, int S, int THREADS>
ZKB_D void ntt_dif_smem_c_from(u64* sm) {       // radix-8 passes from position S down, then the radix-4 / radix-2 remainder
    if constexpr (S >= 3) {
        ntt_pass_c<3, S, L, THREADS>(sm);
        __syncthreads();
        ntt_dif_smem_c_from<L, S - 3, THREADS>(sm);
    } else if constexpr (S > 0) {
        ntt_pass_c<S, S, L, THREADS>(sm);
        __syncthreads();
    }
}

// coset LDE of one column block: out[jb * n + i] = sum_k coeff[k] (shift w_N^j)^k w_n^(k bitrev(i)),  j = bitrev_r(jb)
// prescale: [2^rate_bits][n] table of (shift w_N^j)^k, or null for the plain transform (shift 1, rate 0)
// src_block_stride = 0: every block jb transforms the same n coefficients (LDE); = n: block jb transforms its own
// contiguous run (second step of the two-step transform for n > 2^14; coeffs may alias out). inv: inverse transform.
template <int KMAX, int THREADS, int LGN = 0>      // LGN != 0: forward transform of exactly 2^LGN points, compile-time passes
__global__ void __launch_bounds__(THREADS) lde_block_kernel_t(const u64* coeffs, size_t coeff_stride, u64* out,
                                                              size_t out_stride, unsigned lg_n, const u64* __restrict__ prescale,
                                                              size_t src_block_stride, int inv, unsigned jb0) {
    extern __shared__ u64 sm[];
    const unsigned n = LGN ? (1u << LGN) : (1u << lg_n), jb = blockIdx.x;   // jb: destination block; jb0 + jb: coset (pre-scale table row)
    const u64* src = coeffs + (size_t)blockIdx.y * coeff_stride + (size_t)jb * src_block_stride;
    const u64* ps = prescale ? prescale + (size_t)(jb0 + jb) * n : nullptr;
    // n is a multiple of U * blockDim for the large sizes: U independent loads in flight per thread
    constexpr int U = KMAX == 4 ? 8 : 4;
    if ((n & (U * THREADS - 1)) == 0 && blockDim.x == THREADS) {
#pragma unroll
        for (unsigned i0 = threadIdx.x; i0 < n; i0 += U * THREADS) {
            u64 v[U], w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = src[i0 + u * THREADS];
            if (ps) {
#pragma unroll
                for (int u = 0; u < U; ++u) w[u] = __ldg(ps + i0 + u * THREADS);
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = f_mul(v[u], w[u]);
            }
            // compile-time passes are forward only: the inverse transform is the forward one of the index-negated input
#pragma unroll
            for (int u = 0; u < U; ++u) sm[ntt_pad(LGN && inv ? (n - (i0 + u * THREADS)) & (n - 1) : i0 + u * THREADS)] = v[u];
        }
    } else {
        for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
            u64 v = src[i];
            if (ps) v = f_mul(v, ps[i]);
            sm[ntt_pad(LGN && inv ? (n - i) & (n - 1) : i)] = v;
        }
    }
    __syncthreads();
    if constexpr (LGN != 0) ntt_dif_smem_c_from<LGN, LGN, THREADS>(sm);
    else ntt_dif_smem<KMAX>(sm, lg_n, 0, inv != 0);
    u64* dst = out + (size_t)blockIdx.y * out_stride + (size_t)jb * n;
#pragma unroll 8
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) dst[i] = sm[ntt_pad(i)];
}

// First step of the two-step transform for n = n1 * n2 > 2^14 (view a column as an n1 x n2 matrix, row r = elements
// [r n2, (r+1) n2)): a CTA takes a tile of TB consecutive matrix columns, runs the n1-point DIF down each of them in
// shared memory, multiplies entry (row, b) by w_n^(+-b * bitrev(row)) and writes it back; the rows are then independent
// n2-point transforms (lde_block_kernel). pre1/pre2: optional coset pre-scale factors (s^n2)^r and s^b, s = shift w_N^j.
struct ColsNttArgs {
    const u64* src; size_t src_stride;      // column c, coset-independent source (n elements)
    u64* dst; size_t dst_stride;            // column c, block jb at + jb * n
    unsigned lg_n, lg_n1, lg_tb;
    const u64* pre1; const u64* pre2;       // [2^rate][n1], [2^rate][n2] or null
    int inv;
    unsigned jb0;                           // coset of destination block 0 (rows of pre1 / pre2)
};
__global__ void __launch_bounds__(256) ntt_cols_kernel(ColsNttArgs a) {
    extern __shared__ u64 sm[];
    const unsigned lg_n2 = a.lg_n - a.lg_n1, n1 = 1u << a.lg_n1, n2 = 1u << lg_n2, TB = 1u << a.lg_tb;
    const unsigned b0 = blockIdx.x << a.lg_tb, jb = blockIdx.z;
    const u64* src = a.src + (size_t)blockIdx.y * a.src_stride;
    u64* dst = a.dst + (size_t)blockIdx.y * a.dst_stride + ((size_t)jb << a.lg_n);
    const unsigned tot = n1 << a.lg_tb;
    for (unsigned idx = threadIdx.x; idx < tot; idx += blockDim.x) {
        const unsigned r = idx >> a.lg_tb, c = idx & (TB - 1);
        u64 v = src[(size_t)r * n2 + b0 + c];
        if (a.pre1) v = f_mul(v, f_mul(a.pre1[(size_t)(a.jb0 + jb) * n1 + r], a.pre2[(size_t)(a.jb0 + jb) * n2 + b0 + c]));
        sm[ntt_pad(idx)] = v;
    }
    __syncthreads();
    ntt_dif_smem(sm, a.lg_n1 + a.lg_tb, a.lg_tb, a.inv != 0);
    for (unsigned idx = threadIdx.x; idx < tot; idx += blockDim.x) {
        const unsigned r = idx >> a.lg_tb, c = idx & (TB - 1);
        const unsigned q1 = bitrev32(r, a.lg_n1), b = b0 + c;
        u32 E = (u32)(((u64)b * q1) << (32 - a.lg_n));       // b q1 < n
        if (a.inv) E = 0u - E;
        // T^E from the split tables; for n <= 2^21 the low 11 bits of E are zero, so two factors (and two multiplies) suffice
        u64 v = f_mul(sm[ntt_pad(idx)], d_rootB[(E >> 11) & 2047]);
        v = f_mul(v, d_rootC[E >> 22]);
        if (a.lg_n > 21) v = f_mul(v, d_rootA[E & 2047]);
        dst[(size_t)r * n2 + b] = v;
    }
}

// First step for n = n1 * 2^lg_n2 with n1 <= 64, in registers: ONE THREAD per matrix column b runs the whole n1-point DIF on
// its n1 elements (stride n2 in memory, so a warp's loads and stores are coalesced) — every internal twiddle is a power of two
// (w_64 = 8), no shared memory, no barrier — and multiplies entry (position p, b) by ONE table value
//   tw[jb][p][b] = s_j^b * w_n^(+-b * bitrev(p)),   s_j = shift * w_N^j   (s_j = 1: the plain transform, one block of the table)
// that merges the coset pre-scale's s_j^b with the step twiddle (the generic kernel above spends four multiplies per element
// where this spends two). The thread loops over the coset blocks it writes, re-reading its n1 coefficients (L1/L2 hits: a
// column's coefficients reach HBM once, not once per coset). src may alias dst only when nblk == 1.
struct ColsRegArgs {
    const u64* src; size_t src_stride;
    u64* dst; size_t dst_stride;
    unsigned lg_n2, nblk, jb0;
    const u64* pre1;                        // [2^rate][n1]: (s_j^n2)^r, or null (plain transform)
    const u64* tw; size_t tw_block_stride;  // block jb0 + jb of the table at + (jb0 + jb) * tw_block_stride (0 for the plain one)
    unsigned nsub; size_t sub_stride;       // blockIdx.x = column * nsub + sub: transform number `sub` of a column at + sub * sub_stride
};                                          // (the middle step of a three-step transform; nsub = 1 otherwise)
struct ColsRegBase { const u64* src; u64* dst; };
ZKB_D ColsRegBase cols_reg_base(const ColsRegArgs& a) {
    const unsigned col = blockIdx.x / a.nsub, sub = blockIdx.x - col * a.nsub;
    const size_t b = (size_t)blockIdx.y * 128 + threadIdx.x, off = (size_t)sub * a.sub_stride + b;
    return ColsRegBase{a.src + (size_t)col * a.src_stride + off, a.dst + (size_t)col * a.dst_stride + off};
}
template <int LG1, bool INV>
__global__ void __launch_bounds__(128) ntt_cols_reg_kernel(ColsRegArgs a) {
    constexpr int N1 = 1 << LG1;
    const size_t n2 = size_t(1) << a.lg_n2;
    const size_t b = (size_t)blockIdx.y * 128 + threadIdx.x;
    const ColsRegBase base = cols_reg_base(a);
    const u64* src = base.src;
    u64* dst0 = base.dst;
#pragma unroll 1
    for (unsigned jb = 0; jb < a.nblk; ++jb) {
        u64 r[N1];
#pragma unroll
        for (int e = 0; e < N1; ++e) r[e] = src[(size_t)e * n2];
        if (a.pre1) {
            const u64* p1 = a.pre1 + (size_t)(a.jb0 + jb) * N1;
#pragma unroll
            for (int e = 1; e < N1; ++e) r[e] = f_mul(r[e], __ldg(p1 + e));
        }
        RadixStep<LG1, INV, 0, 0, 0>::run(r);
        const u64* tw = a.tw + (size_t)(a.jb0 + jb) * a.tw_block_stride + b;
        u64* dst = dst0 + ((size_t)jb << (LG1 + a.lg_n2));
        if (a.pre1) r[0] = f_mul(r[0], __ldg(tw));
        dst[0] = r[0];
#pragma unroll
        for (int p = 1; p < N1; ++p) dst[(size_t)p * n2] = f_mul(r[p], __ldg(tw + (size_t)p * n2));
    }
}
// The same first step for n1 = 32 / 64 = 8 x NB, where n1 values per thread no longer fit in registers (the 64-point form of
// the kernel above needs 255 registers, runs 4 warps per scheduler... 12 % occupancy, 56 % instruction-cache hits: 3x slower per
// element than the 16-point one — profiles/r02_ncu_cols_reg.md). Each thread stages its column in a PRIVATE strip of shared
// memory (word [row][threadIdx.x]: conflict-free, no barrier):
//   phase 1, for each residue a < 8 (unrolled, so every twiddle stays a compile-time shift): the NB-point DIF over e of the
//            pre-scaled x[a + 8 e], times w_n1^(a q), q = the frequency left at position p;
//   phase 2, for each p < NB (one rolled loop): the 8-point DIF over a, the merged table multiply, the store to row 8 p + p2
//            — the position the plain n1-point DIF would have left frequency bitrev(8 p + p2) in.
__host__ __device__ constexpr int bitrev_c(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
template <int LGB, int A, int MUL, bool INV, int... P>
ZKB_D void staged_twiddle(u64* r, std::integer_sequence<int, P...>) {
    ((r[P] = f_mul_w64<(INV ? -1 : 1) * A * bitrev_c(P, LGB) * MUL>(r[P])), ...);
}
template <int LG1, bool INV, int A>
ZKB_D void staged_phase1(const u64* src, size_t n2, const u64* p1, u64* my) {
    constexpr int LGB = LG1 - 3, NB = 1 << LGB;
    u64 r[NB];
#pragma unroll
    for (int e = 0; e < NB; ++e) r[e] = src[(size_t)(A + 8 * e) * n2];
    if (p1) {
#pragma unroll
        for (int e = 0; e < NB; ++e)
            if (A + 8 * e) r[e] = f_mul(r[e], __ldg(p1 + A + 8 * e));
    }
    RadixStep<LGB, INV, 0, 0, 0>::run(r);
    staged_twiddle<LGB, A, (64 >> LG1), INV>(r, std::make_integer_sequence<int, NB>{});
#pragma unroll
    for (int p = 0; p < NB; ++p) my[(A + 8 * p) * 128] = r[p];
}
template <int LG1, bool INV, int... A>
ZKB_D void staged_phase1_all(const u64* src, size_t n2, const u64* p1, u64* my, std::integer_sequence<int, A...>) {
    (staged_phase1<LG1, INV, A>(src, n2, p1, my), ...);
}
template <int LG1, bool INV>
__global__ void __launch_bounds__(128) ntt_cols_staged_kernel(ColsRegArgs a) {
    constexpr int NB = 1 << (LG1 - 3);
    extern __shared__ u64 sm[];             // [2^LG1][128]
    u64* my = sm + threadIdx.x;
    const size_t n2 = size_t(1) << a.lg_n2;
    const size_t b = (size_t)blockIdx.y * 128 + threadIdx.x;
    const ColsRegBase base = cols_reg_base(a);
    const u64* src = base.src;
    u64* dst0 = base.dst;
#pragma unroll 1
    for (unsigned jb = 0; jb < a.nblk; ++jb) {
        const u64* p1 = a.pre1 ? a.pre1 + ((size_t)(a.jb0 + jb) << LG1) : nullptr;
        staged_phase1_all<LG1, INV>(src, n2, p1, my, std::make_integer_sequence<int, 8>{});
        const u64* tw = a.tw + (size_t)(a.jb0 + jb) * a.tw_block_stride + b;
        u64* dst = dst0 + ((size_t)jb << (LG1 + a.lg_n2));
#pragma unroll 1
        for (int p = 0; p < NB; ++p) {
            u64 r[8], w[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = __ldg(tw + (size_t)(8 * p + k) * n2);
#pragma unroll
            for (int k = 0; k < 8; ++k) r[k] = my[(k + 8 * p) * 128];
            RadixStep<3, INV, 0, 0, 0>::run(r);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                u64 v = r[k];
                if (a.pre1 || p + k) v = f_mul(v, w[k]);
                dst[(size_t)(8 * p + k) * n2] = v;
            }
        }
    }
}
// tw[jb][p][b] of the kernel above; base_j = shift * w_N^bitrev_r(jb) (1 for the plain table), w = w_n or its inverse
__global__ void cols_reg_table_kernel(u64* tw, unsigned lg_n1, unsigned lg_n2, unsigned rate_bits, u64 shift, u64 w_N, u64 w_n, int coset) {
    const size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const unsigned p = blockIdx.y, jb = blockIdx.z;
    if (b >= (size_t(1) << lg_n2)) return;
    u64 base = gl_pow(w_n, bitrev32(p, lg_n1));
    if (coset) base = gl_mul(base, gl_mul(shift, gl_pow(w_N, bitrev32(jb, rate_bits))));
    tw[((((size_t)jb << lg_n1) + p) << lg_n2) + b] = gl_pow(base, b);
}

// inverse transform of one column: values on <w_n> (natural order, or bit-reversed if in_bitrev) -> coefficients
// (natural order), times `scale` (= 1/n). Uses the forward code: c_k = X[(n - k) mod n] / n.
// the same with compile-time passes for the circuit sizes (radix-8, 1024 / 512 threads)
template <int LGN, int THREADS>
__global__ void __launch_bounds__(THREADS) intt_block_kernel_c(const u64* __restrict__ in, size_t in_stride, u64* __restrict__ out,
                                                               size_t out_stride, int in_bitrev, u64 scale) {
    extern __shared__ u64 sm[];
    constexpr unsigned n = 1u << LGN;
    const u64* src = in + (size_t)blockIdx.x * in_stride;
#pragma unroll 4
    for (unsigned i = threadIdx.x; i < n; i += THREADS) {
        const unsigned pos = in_bitrev ? bitrev32(i, LGN) : i;
        sm[ntt_pad(pos)] = src[i];
    }
    __syncthreads();
    ntt_dif_smem_c_from<LGN, LGN, THREADS>(sm);
    u64* dst = out + (size_t)blockIdx.x * out_stride;
#pragma unroll 4
    for (unsigned k = threadIdx.x; k < n; k += THREADS) {
        const unsigned q = (n - k) & (n - 1);
        dst[k] = f_mul(sm[ntt_pad(bitrev32(q, LGN))], scale);
    }
}
__global__ void __launch_bounds__(512) intt_block_kernel(const u64* __restrict__ in, size_t in_stride, u64* __restrict__ out,
                                                         size_t out_stride, unsigned lg_n, int in_bitrev, u64 scale) {
    extern __shared__ u64 sm[];
    const unsigned n = 1u << lg_n;
    const u64* src = in + (size_t)blockIdx.x * in_stride;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned pos = in_bitrev ? bitrev32(i, lg_n) : i;
        sm[ntt_pad(pos)] = src[i];
    }
    __syncthreads();
    ntt_dif_smem(sm, lg_n);
    u64* dst = out + (size_t)blockIdx.x * out_stride;
    for (unsigned k = threadIdx.x; k < n; k += blockDim.x) {
        const unsigned q = (n - k) & (n - 1);
        dst[k] = gl_mul(sm[ntt_pad(bitrev32(q, lg_n))], scale);
    }
}

// table[jb * n + k] = (shift * w_N^bitrev_r(jb))^k, N = n << rate_bits; built once per (n, rate, shift)
__global__ void coset_table_kernel(u64* table, unsigned lg_n, unsigned rate_bits, u64 shift, u64 w_N) {
    const unsigned n = 1u << lg_n;
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x, jb = blockIdx.y;
    if (k >= n) return;
    const u64 base = gl_mul(shift, gl_pow(w_N, bitrev32(jb, rate_bits)));
    table[(size_t)jb * n + k] = gl_pow(base, k);
}

}  // namespace zkb
