// Host-callable launchers for the sm_100a kernels in kernels.cu. Everything here takes DEVICE
// pointers and a stream; nothing synchronises. Data layout conventions:
//   * polynomial batches are column-major: element (col, idx) at base[col * stride + idx]
//   * LDE values are stored in the reference's leaf order: position l of a column holds the value at
//     natural LDE index bitrev(l)  (qp-plonky2 `reverse_index_bits_in_place` on the leaves)
//   * digests are 4 x u64, array-of-structs; a tree's levels are concatenated bottom-up
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include "field.cuh"
#include "poseidon_consts.hpp"   // host_round_constants()

namespace zkb {

// ---- one-time init (per device): Poseidon constants + twiddle tables ----
void device_tables_init(int device);            // idempotent, thread-safe
unsigned long long kernel_launch_count();       // kernels launched by this library so far (process-wide)
void kernel_launch_count_add(unsigned long long n);   // replayed CUDA-graph kernel nodes

// ---- Poseidon / Merkle ----
void launch_poseidon_permute(u64* states, size_t count, cudaStream_t st);
// leaves[c * col_stride + l], c < width, l < num_leaves  ->  digests[l*4 .. l*4+3]
void launch_merkle_leaves(const u64* leaves, size_t col_stride, int width, size_t num_leaves, u64* digests, cudaStream_t st);
// FRI layer leaves: leaf l = 2*arity felts (v[arity*l + k].a, v[arity*l + k].b) from SoA arrays a[], b[]
void launch_merkle_leaves_ext(const u64* a, const u64* b, int arity, size_t num_leaves, u64* digests, cudaStream_t st);
// builds all levels above level 0 in `digests` (levels concatenated: num_leaves, num_leaves/2, ...) down to
// 2^cap_height nodes; returns the offset (in digests) of the cap level
size_t launch_merkle_levels(u64* digests, size_t num_leaves, unsigned cap_height, cudaStream_t st);
// whole trees (leaf digests + every level, 2 launches for the prover's trees): return the digest offset of the cap level
size_t launch_merkle_tree(const u64* leaves, size_t col_stride, int width, size_t num_leaves, u64* digests, unsigned cap_height,
                          cudaStream_t st);
size_t launch_merkle_tree_ext(const u64* a, const u64* b, int arity, size_t num_leaves, u64* digests, unsigned cap_height,
                              cudaStream_t st);
size_t merkle_digest_count(size_t num_leaves, unsigned cap_height);   // total digests over all levels
size_t merkle_level_offset(size_t num_leaves, unsigned level);        // digest offset of level k

// ---- NTT family ----
// values (natural order) -> coefficients (natural order), ncols columns of size n = 2^lg_n, in place allowed
void launch_intt_natural(const u64* src, size_t src_stride, u64* dst, size_t dst_stride, int ncols, unsigned lg_n,
                         u64* scratch /* ncols*n words if lg_n > max smem block, else may be null */, cudaStream_t st);
// coefficients (natural, n per column) -> evaluations on shift*<w_{n<<rate_bits}> in leaf (bit-reversed) order
void launch_lde(const u64* coeffs, size_t coeff_stride, u64* out, size_t out_stride, int ncols, unsigned lg_n,
                unsigned rate_bits, u64 shift, cudaStream_t st);
// the same for the leaf blocks [blk_lo, blk_hi) only (block jb = LDE coset bitrev(jb) = leaves [jb n, (jb+1) n));
// out points at the destination of block blk_lo (coset sharding across GPUs)
void launch_lde_blocks(const u64* coeffs, size_t coeff_stride, u64* out, size_t out_stride, int ncols, unsigned lg_n,
                       unsigned rate_bits, u64 shift, unsigned blk_lo, unsigned blk_hi, cudaStream_t st);
// evaluations on shift*<w_m> given in bit-reversed order -> coefficients in natural order (in place)
void launch_coset_intt_bitrev(u64* data, size_t stride, int ncols, unsigned lg_m, u64 shift, cudaStream_t st);

// out[m * len + k] = sum_j mat[m * R + j] * in[j * len + k], R <= 16 (chunk recovery of the coset-sharded quotient)
void launch_vandermonde_solve(const u64* in, u64* out, size_t len, unsigned R, const u64* mat_dev, cudaStream_t st);

// in-place bit-reversal permutation of each column (leaf order <-> natural order)
void launch_bitrev_permute(u64* data, size_t stride, int ncols, unsigned lg_n, cudaStream_t st);

// *flag_dev |= 1 if any of the count elements is not a canonical field element
void launch_canonical_check(const u64* v, size_t count, unsigned* flag_dev, cudaStream_t st);
// salt columns: out[s * stride + l] = salt_value(seed, batch, s, l)  (documented SplitMix64 generator)
void launch_salt_fill(u64* out, size_t stride, size_t num_leaves, u64 seed, unsigned batch, cudaStream_t st);
// production salts: `words` field elements from the ChaCha20 key stream of `key` (256 bits from the OS RNG, per proof),
// nonce = batch; out is the flat [4][stride] salt block of a batch
void launch_salt_fill_csprng(u64* out, size_t words, const u32 key[8], unsigned batch, cudaStream_t st);

// ---- prover stages ----
// param / p2 / p3: Constant num_consts; BaseSum num_limbs; Arithmetic, ArithmeticExtension, MulExtension num_ops;
// Reducing(Extension) num_coeffs; Exponentiation num_power_bits; RandomAccess bits / num_copies / num_extra_constants;
// CosetInterpolation subgroup_bits / degree
struct GateDesc { u32 tag; u32 param; u32 selector_index; u32 group_lo, group_hi; u32 row; u32 p2, p3; };
struct QuotientParams {
    unsigned lg_n, rate_bits;
    int num_wires, num_routed, num_constants, num_selectors, num_challenges, num_partial_products, qdf;
    int num_gates, num_gate_constraints;
    GateDesc gates[32];
    u64 k_is[128];
    u64 betas[4], gammas[4], alphas[4];
    u64 pi_hash[4];
    u64 zh_inv[16];      // 1 / Z_H on the coset, index i mod 2^rate_bits
    u64 zh[16];
    u64 n_inv_dummy;
    u64 bary_w[16], bary_x[16];   // CosetInterpolation: barycentric weights and the subgroup points w^k
    int has_recursion_gates;      // any gate evaluated by the third quotient launch
};
// chunk products + running product -> out columns [Z_0..Z_{c-1}, pp(ch0)..., pp(ch1)...], each n values
// betas_gammas: HOST array [betas(nch), gammas(nch)]
void launch_partial_products(const u64* wires, size_t wire_stride, const u64* sigma_values, size_t sigma_stride,
                             const u64* k_is_dev, int num_routed, int chunk, int num_challenges, const u64* betas_gammas,
                             unsigned lg_n, u64* out, size_t out_stride, u64* scratch, cudaStream_t st);
size_t partial_products_scratch_words(int num_routed, int chunk, int num_challenges, unsigned lg_n);
// quotient values at every LDE point (leaf order): out[ch * out_stride + l]
// apow_dev: [num_challenges][nterms] powers of alpha, nterms = nch*(2+npp) + num_gate_constraints
// fork: null for the single-stream form (large circuits); otherwise the three launches run concurrently on st and the two
// helper streams with their gates spread over blockIdx.y, into `part` = quotient_slots() x num_challenges x N words
struct QuotientFork { u64* part; cudaStream_t aux[2]; cudaEvent_t fork, join[2]; };
int quotient_slots(const QuotientParams& params_host);
void launch_quotient(const QuotientParams* params_dev, const QuotientParams& params_host, const u64* apow_dev, int nterms,
                     const u64* cs_lde, size_t cs_stride, const u64* wires_lde, size_t w_stride, const u64* zs_lde,
                     size_t z_stride, u64* out, size_t out_stride, cudaStream_t st, const QuotientFork* fork = nullptr);
// ZKB_CHECK_WITNESS: evaluate every constraint (permutation argument included) at the n points of the subgroup H from VALUES
// in natural order (cs_vals [num_constants + num_routed][n], wires_vals, zs_vals) with the powers of a check challenge in
// apow_dev; *flag_dev |= 2 if any point's combination is non-zero. scratch: num_challenges * n words.
void launch_constraint_check(const QuotientParams* params_dev, const QuotientParams& params_host, const u64* apow_dev, int nterms,
                             const u64* cs_vals, size_t cs_stride, const u64* wires_vals, size_t w_stride, const u64* zs_vals,
                             size_t z_stride, u64* scratch, size_t scratch_stride, unsigned* flag_dev, cudaStream_t st);
// evaluate ncols coefficient polynomials (n each) at the ext point z: out[2*c], out[2*c+1]
void launch_eval_polys(const u64* coeffs, size_t stride, int ncols, unsigned lg_n, const u64* zpow_a, const u64* zpow_b,
                       u64* out, cudaStream_t st);
// every opening of a proof: out[2 i], out[2 i + 1] = P_i(point) for the polynomials of the segments in order.
// pw: scratch of 4 n words (powers of the two points); point 0 = z0, 1 = z1
struct OpeningsSeg { const u64* coeffs; size_t stride; int ncols; int point; };
struct OpeningsArgs { OpeningsSeg seg[6]; int nseg; unsigned lg_n; u64* pw; u64* out; };
void launch_openings(const OpeningsArgs& a, ext2 z0, ext2 z1, cudaStream_t st);
// zpow[k] = z^k for k < n (SoA)
void launch_ext_powers(ext2 z, unsigned lg_n, u64* zpow_a, u64* zpow_b, cudaStream_t st);

struct FriCombineParams {
    const u64* lde[4]; size_t stride[4]; int ncols[4];   // unsalted column counts per committed batch
    int num_zs;                                          // columns of batch 2 (Z / partial products) opened at g*zeta
    ext2 alpha, zeta, zeta_next, reduced0, reduced1;
    unsigned lg_n;
};
// q(x) on the coset g*H (first n leaves), leaf order, SoA out_a/out_b
void launch_fri_combine(const FriCombineParams& p, const u64* alpha_pows_a, const u64* alpha_pows_b, u64* out_a, u64* out_b, cudaStream_t st);
// coefficient fold: out[k] = sum_{i<arity} beta^i c[arity*k + i], k < m_out
void launch_fri_fold(const u64* ca, const u64* cb, u64* oa, u64* ob, size_t m_out, int arity, ext2 beta, cudaStream_t st);
// proof-of-work: smallest w in [base, base+count) with leading_zeros(permute(state with state[pos]=w)[7]) >= bits;
// *result (device) must be preset to ~0ull
void launch_pow_search(const u64* state12_dev, int pos, u64 base, u64 count, unsigned bits, unsigned long long* result, cudaStream_t st);
// gather rows: out[q * width + c] = lde[c * stride + idx[q]]
void launch_gather_rows(const u64* lde, size_t stride, int width, const u32* idx_dev, int nq, u64* out, cudaStream_t st);
// gather merkle paths: out[(q * path_len + k) * 4 ..] = digests[level k][(idx[q] >> k) ^ 1]
void launch_gather_paths(const u64* digests, size_t num_leaves, int path_len, const u32* idx_dev, int nq, u64* out, cudaStream_t st);
// gather FRI layer leaves from SoA: out[(q*arity + k)*2 + {0,1}] = (a,b)[arity*idx[q] + k]
void launch_gather_ext_leaves(const u64* a, const u64* b, int arity, const u32* idx_dev, int nq, u64* out, cudaStream_t st);

}  // namespace zkb
