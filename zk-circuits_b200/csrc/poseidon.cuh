// Width-12 Goldilocks Poseidon (4 + 22 + 4 rounds, x^7) for sm_100a — the hash behind
// `PoseidonGoldilocksConfig` (/root/reference/common/src/circuit.rs:10-12): leaf hashing, 2-to-1
// compression, the Fiat-Shamir duplex and the proof-of-work grind of qp-plonky2 1.1.1.
//
// Design. Measured on B200 (profiles/r01_ubench_int_pipes.md): IMAD.WIDE costs two issue slots
// and does not overlap with the carry adds around it, while 32-bit IADD3/LEA/IMAD issue on two independent pipes.
// So the LINEAR layer never touches a wide multiply:
//   * a state word lives as three signed 32-bit limbs, value = l0 + l1*2^22 + l2*2^44 (mod p);
//   * the MDS matrix circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,..) is applied to each limb vector as a
//     length-12 cyclic correlation split 12 = 3 x 4: a 4-point DFT over Z[i] of the three residue classes, a 3x3
//     twisted product whose frequency-domain constants are all powers of two ({16,16,32}, {-1,-2,8},
//     {2+i, 1+16i, 1-4i}), and the inverse DFT with the 1/4 folded into those constants — 68 shift/add
//     instructions per limb vector instead of 144 multiply-adds, exact in wrap-around 32-bit arithmetic because the
//     true outputs stay inside (-2^31, 2^31);
//   * words that skip the S-box (11 of 12 in the 22 partial rounds) are only carry-normalised (10 instructions);
//     words that take the S-box are folded to one lazy 64-bit value, raised to the 7th power with 4 full 64x64
//     multiplies (field.cuh), and split again.
// Round constants (as limbs) come from __constant__ memory; constants of round r are added to the limbs at the
// start of round r, so a sponge can overwrite rate words between permutations without leaving limb form.
// ONE loop over the 30 rounds (warp-uniform full/partial switch, S-boxes in 3 passes of 4 with a register
// rotation) keeps the code inside the SM's instruction cache.
#pragma once
#include "field.cuh"

namespace zkb {

constexpr int P_WIDTH = 12;
constexpr int P_RATE = 8;
constexpr int P_HALF_FULL = 4;
constexpr int P_PARTIAL = 22;
constexpr int P_ROUNDS = 30;
constexpr u32 P_M22 = (1u << 22) - 1, P_M20 = (1u << 20) - 1;

// host-only self-check of the range argument behind the 32-bit limb arithmetic (tests/native build with -DZKB_CHECK_BOUNDS)
#if defined(ZKB_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
#include <cstdlib>
#define ZKB_BOUND_CHECK(cond) do { if (!(cond)) { std::abort(); } } while (0)
#else
#define ZKB_BOUND_CHECK(cond) do { } while (0)
#endif

ZKB_HD u64 gl_sbox7(u64 x) {
    u64 x2 = gl_sqr_lazy(x);
    u64 x3 = gl_mul_lazy(x2, x);
    u64 x4 = gl_sqr_lazy(x2);
    return gl_mul_lazy(x3, x4);
}

// ---- limb form -------------------------------------------------------------------------------
// "raw" limbs: any signed 32-bit values with |l| < 2^31 - 2^22; "normalised": l0 in (-2^11, 2^22 + 2^11],
// l1 in [-2^21, 2^22 + 2^21), l2 in [0, 2^20). With normalised inputs every MDS output is below
// 272 * (2^22 + 2^21) < 2^31 - 2^22 in magnitude, so adding a round-constant limb (< 2^22) cannot overflow.
ZKB_HD void limb_split(u64 v, u32& l0, u32& l1, u32& l2) {
    u32 lo = (u32)v, hi = (u32)(v >> 32);
    l0 = lo & P_M22;
    l1 = ((lo >> 22) | (hi << 10)) & P_M22;
    l2 = hi >> 12;
}
ZKB_HD void limb_normalize(u32& x0, u32& x1, u32& x2) {
    int c0 = (int)x0 >> 22;
    x0 &= P_M22;
    x1 += (u32)c0;
    int c1 = (int)x1 >> 22;
    x1 &= P_M22;
    x2 += (u32)c1;
    int top = (int)x2 >> 20;                 // multiples of 2^64 = 2^32 - 1 (mod p)
    x2 &= P_M20;
    x1 += (u32)top << 10;
    x0 -= (u32)top;
}
// raw limbs -> one lazy field element in [0, 2^64)
ZKB_HD u64 limb_to_u64(u32 x0, u32 x1, u32 x2) {
    int c0 = (int)x0 >> 22;
    x0 &= P_M22;
    x1 += (u32)c0;
    int c1 = (int)x1 >> 22;
    x1 &= P_M22;
    x2 += (u32)c1;
    int top = (int)x2 >> 20;
    x2 &= P_M20;
    u32 lo = x0 | (x1 << 22), hi = (x1 >> 10) | (x2 << 12);
    u64 w = ((u64)hi << 32) | lo;
    if (top >= 0) {                          // + top * 2^64
        u64 k = (u64)(u32)top * GL_EPS;      // < 2^43
        u64 r = w + k;
        if (r < k) r += GL_EPS;              // wrapped once; r < 2^43 so this cannot wrap again
        return r;
    }
    u64 k = (u64)(u32)(-top) * GL_EPS;       // rare: the top limb went negative through a borrow
    u64 r = w - k;
    if (w < k) r -= GL_EPS;
    return r;
}

// Same, for raw limbs that carry the round-constant bias (see host_round_constant_limbs): the top carry is >= 0 by
// construction, so there is no sign case and no branch.
ZKB_HD u64 limb_to_u64_biased(u32 x0, u32 x1, u32 x2) {
    int c0 = (int)x0 >> 22;
    x0 &= P_M22;
    x1 += (u32)c0;
    int c1 = (int)x1 >> 22;
    x1 &= P_M22;
    x2 += (u32)c1;
    ZKB_BOUND_CHECK((int)x2 >= 0);           // the biased constants keep the top limb non-negative
    u32 top = x2 >> 20;                      // >= 1: the bias puts 2^20 into limb 2
    x2 &= P_M20;
    u32 lo = x0 | (x1 << 22), hi = (x1 >> 10) | (x2 << 12);
    u64 w = ((u64)hi << 32) | lo;
    u64 k = ((u64)top << 32) - top;          // top * (2^32 - 1) < 2^44
    u64 r = w + k;
    if (r < k) r += GL_EPS;
    return r;
}

// o = circ(17,15,41,16,2,28,13,13,39,18,34,20) x + 8 x[0] e_0 on one limb vector, arithmetic mod 2^32
// (out[r] = sum_i x[(i+r) % 12] C[i]); derivation and the Python check of these formulas: DESIGN.md §4.1.
// rc: limb k of the next round's 12 constants at rc[3 * j], added into the output sums (free third IADD3 operand)
ZKB_HD void mds_limb12(u32* x, const u32* __restrict__ rc) {
#if defined(ZKB_CHECK_BOUNDS) && !defined(__CUDA_ARCH__)
    for (int j = 0; j < 12; ++j) ZKB_BOUND_CHECK((int)x[j] > -(1 << 21) - 1 && (int)x[j] < (1 << 22) + (1 << 21));   // normalised inputs
#endif
    u32 P[3], M[3], R[3], I[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        u32 t0 = x[b] + x[6 + b], t1 = x[3 + b] + x[9 + b];
        P[b] = t0 + t1;
        M[b] = t0 - t1;
        R[b] = x[b] - x[6 + b];
        I[b] = x[3 + b] - x[9 + b];
    }
    const u32 x0 = x[0];
    const u32 T = P[0] + P[1] + P[2];
    u32 G[3] = {T + P[2], T + P[0], T + P[1]};                       // (k = 0 component) / 16
    u32 F[3] = {8 * M[2] - M[0] - 2 * M[1], 0u - M[1] - 2 * M[2] - 8 * M[0], 2 * M[0] - M[2] - 8 * M[1]};   // k = 2
    u32 Ar[3], Ai[3], Br[3], Bi[3], Dr[3], Di[3];                    // Z_b * (2+i), * (1+16i), * (1-4i), Z_b = R_b + i I_b
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        Ar[b] = 2 * R[b] - I[b];  Ai[b] = R[b] + 2 * I[b];
        Br[b] = R[b] - 16 * I[b]; Bi[b] = 16 * R[b] + I[b];
        Dr[b] = R[b] + 4 * I[b];  Di[b] = I[b] - 4 * R[b];
    }
    u32 Re[3] = {Ar[0] + Br[1] + Dr[2], Ar[1] + Br[2] + Di[0], Ar[2] + Bi[0] + Di[1]};
    u32 Im[3] = {Ai[0] + Bi[1] + Di[2], Ai[1] + Bi[2] - Dr[0], Ai[2] - Br[0] - Dr[1]};
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        u32 U = 16 * G[b] + F[b], V = 16 * G[b] - F[b];
        x[b] = U + Re[b] + rc[3 * b];
        x[6 + b] = U - Re[b] + rc[3 * (6 + b)];
        x[3 + b] = V + Im[b] + rc[3 * (3 + b)];
        x[9 + b] = V - Im[b] + rc[3 * (9 + b)];
    }
    x[0] += 8 * x0;
}

// The permutation on limb-form state. rc3[36 * r + 3 * j + k] = limb k of the BIASED round constant (r, j), r = 0..30:
// value = rc(r, j) for r < 30 and 0 for r = 30, written as limbs((value - (2^32 - 1)) mod p) + 2^20 in limb 2, which is
// the same field element (2^64 = 2^32 - 1 mod p) but keeps every top carry non-negative.
// In: raw limbs WITHOUT the first round's constants. Out: raw limbs of the last MDS layer plus the biased zero.
ZKB_HD void poseidon_permute_limbs(u32* o0, u32* o1, u32* o2, const u32* __restrict__ rc3) {
#pragma unroll
    for (int j = 0; j < 12; ++j) { o0[j] += rc3[3 * j]; o1[j] += rc3[3 * j + 1]; o2[j] += rc3[3 * j + 2]; }
#pragma unroll 1
    for (int r = 0; r < P_ROUNDS; ++r) {
        if (r < P_HALF_FULL || r >= P_HALF_FULL + P_PARTIAL) {
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    u64 v = gl_sbox7(limb_to_u64_biased(o0[i], o1[i], o2[i]));
                    limb_split(v, o0[i], o1[i], o2[i]);
                }
                u32 t;
#pragma unroll
                for (int i = 0; i < 4; ++i) {   // rotate the three limb vectors left by 4 words
                    t = o0[i]; o0[i] = o0[i + 4]; o0[i + 4] = o0[i + 8]; o0[i + 8] = t;
                    t = o1[i]; o1[i] = o1[i + 4]; o1[i + 4] = o1[i + 8]; o1[i + 8] = t;
                    t = o2[i]; o2[i] = o2[i + 4]; o2[i + 4] = o2[i + 8]; o2[i + 8] = t;
                }
            }
        } else {
            u64 v = gl_sbox7(limb_to_u64_biased(o0[0], o1[0], o2[0]));
            limb_split(v, o0[0], o1[0], o2[0]);
#pragma unroll
            for (int j = 1; j < 12; ++j) limb_normalize(o0[j], o1[j], o2[j]);
        }
        const u32* rc = rc3 + 36 * (r + 1);
        mds_limb12(o0, rc);
        mds_limb12(o1, rc + 1);
        mds_limb12(o2, rc + 2);
    }
}

#if defined(__CUDACC__)
__constant__ u64 c_rc[P_WIDTH * P_ROUNDS];                 // round constants as field elements (quotient kernel)
__constant__ u32 c_rc3[3 * P_WIDTH * (P_ROUNDS + 1)];      // biased limb form, 31 rows (see poseidon_permute_limbs)
__constant__ u64 c_rc2[2 * P_WIDTH * (P_ROUNDS + 1)];      // split 32-bit halves, + 24 zeros (see mds_layer_rc)

// sponge state held in registers in limb form
struct PoseidonState {
    u32 o0[12], o1[12], o2[12];
    ZKB_D void zero() {
#pragma unroll
        for (int j = 0; j < 12; ++j) o0[j] = o1[j] = o2[j] = 0;
    }
    ZKB_D void set(int j, u64 v) { limb_split(v, o0[j], o1[j], o2[j]); }      // v < 2^64 (lazy allowed)
    ZKB_D u64 get(int j) const { return gl_canon(limb_to_u64(o0[j], o1[j], o2[j])); }   // general form: valid before and after permute()
    ZKB_D void permute() { poseidon_permute_limbs(o0, o1, o2, c_rc3); }
};

// u64 in / lazy u64 out wrapper
ZKB_D void poseidon_permute(u64* s) {
    PoseidonState st;
#pragma unroll
    for (int j = 0; j < 12; ++j) st.set(j, s[j]);
    st.permute();
#pragma unroll
    for (int j = 0; j < 12; ++j) s[j] = limb_to_u64(st.o0[j], st.o1[j], st.o2[j]);
}

// ---- 64-bit-word MDS used by the quotient kernel's Poseidon gate (it needs every intermediate state) ----
#define ZKB_MDS_C(i) ((i) == 0 ? 17u : (i) == 1 ? 15u : (i) == 2 ? 41u : (i) == 3 ? 16u : (i) == 4 ? 2u : (i) == 5 ? 28u : \
                      (i) == 6 ? 13u : (i) == 7 ? 13u : (i) == 8 ? 39u : (i) == 9 ? 18u : (i) == 10 ? 34u : 20u)

// out[r] = sum_i s[(i+r)%12] * C[i] + (r==0 ? 8*s[0] : 0) + rc[r]; inputs/outputs lazy u64.
// rc2 points at the split constants of the round whose constants are folded in: rc2[2r] = low 32 bits of rc[r],
// rc2[2r+1] = high 32 bits (each zero-extended to 64 bits so that they initialise the two IMAD.WIDE accumulators).
ZKB_D void mds_layer_rc(u64* s, const u64* rc2) {
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { lo[i] = (u32)s[i]; hi[i] = (u32)(s[i] >> 32); }
#pragma unroll
    for (int r = 0; r < 12; ++r) {
        u64 al = rc2[2 * r], ah = rc2[2 * r + 1];
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            al += (u64)lo[(i + r) % 12] * ZKB_MDS_C(i);
            ah += (u64)hi[(i + r) % 12] * ZKB_MDS_C(i);
        }
        if (r == 0) { al += (u64)lo[0] * 8u; ah += (u64)hi[0] * 8u; }
        u32 b0 = (u32)ah, b1 = (u32)(ah >> 32);                 // b1 < 2^12
        u64 t = (u64)b1 * 0xFFFFFFFFu + al;                      // < 2^45, no overflow
        u32 q0, q1;
        asm("{\n\t.reg .u32 m;\n\t"
            "add.cc.u32 %1, %3, %4;\n\t"      // high limb + b0 -> carry (rare)
            "addc.u32 m, 0, 0;\n\t"
            "neg.s32 m, m;\n\t"
            "add.cc.u32 %0, %2, m;\n\t"       // + EPS on carry (cannot wrap again: the wrapped high limb is < 2^13)
            "addc.u32 %1, %1, 0;\n\t}"
            : "=&r"(q0), "=&r"(q1) : "r"((u32)t), "r"((u32)(t >> 32)), "r"(b0));
        s[r] = ((u64)q1 << 32) | q0;
    }
}
ZKB_D void mds_layer(u64* s) { mds_layer_rc(s, c_rc2 + 24 * P_ROUNDS); }   // the all-zero tail of the table
#endif  // __CUDACC__

}  // namespace zkb
