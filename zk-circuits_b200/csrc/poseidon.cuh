// Width-12 Goldilocks Poseidon (4 + 22 + 4 rounds, x^7) for sm_100a — the hash behind
// `PoseidonGoldilocksConfig` (/root/reference/common/src/circuit.rs:10-12): leaf hashing, 2-to-1
// compression, the Fiat-Shamir duplex and the proof-of-work grind of qp-plonky2 1.1.1.
// Round constants live in __constant__ memory (uploaded once by zkb::poseidon_init); the MDS layer
// works on 32-bit halves with small-constant IMAD.WIDE accumulation (coefficients <= 41, sums < 2^42)
// and a single reduction per output word.
#pragma once
#include "field.cuh"

namespace zkb {

constexpr int P_WIDTH = 12;
constexpr int P_RATE = 8;
constexpr int P_HALF_FULL = 4;
constexpr int P_PARTIAL = 22;
constexpr int P_ROUNDS = 30;

#if defined(__CUDACC__)
__constant__ u64 c_rc[P_WIDTH * P_ROUNDS];
__constant__ u64 c_rc2[2 * P_WIDTH * (P_ROUNDS + 1)];   // split halves, + 24 zeros (see mds_layer_rc)

#define ZKB_MDS_C(i) ((i) == 0 ? 17u : (i) == 1 ? 15u : (i) == 2 ? 41u : (i) == 3 ? 16u : (i) == 4 ? 2u : (i) == 5 ? 28u : \
                      (i) == 6 ? 13u : (i) == 7 ? 13u : (i) == 8 ? 39u : (i) == 9 ? 18u : (i) == 10 ? 34u : 20u)

ZKB_D u64 gl_sbox7(u64 x) {
    u64 x2 = gl_mul_lazy(x, x);
    u64 x3 = gl_mul_lazy(x2, x);
    u64 x4 = gl_mul_lazy(x2, x2);
    return gl_mul_lazy(x3, x4);
}

// out[r] = sum_i s[(i+r)%12] * C[i] + (r==0 ? 8*s[0] : 0) + rc[r]; inputs/outputs lazy u64.
// rc2 points at the split constants of the round whose constants are folded in: rc2[2r] = low 32 bits of rc[r],
// rc2[2r+1] = high 32 bits (each zero-extended to 64 bits so that they initialise the two IMAD.WIDE accumulators).
// Per output: 24 (26) IMAD.WIDE.U32 small-constant MACs on the 32-bit halves (sums < 2^44), then
// al + ah*2^32 = (al + b1*EPS) + b0*2^32 with one IMAD.WIDE and a single-carry fix-up (5 ALU instructions).
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
        u32 o0, o1;
        asm("{\n\t.reg .u32 m;\n\t"
            "add.cc.u32 %1, %3, %4;\n\t"      // high limb + b0 -> carry (rare)
            "addc.u32 m, 0, 0;\n\t"
            "neg.s32 m, m;\n\t"
            "add.cc.u32 %0, %2, m;\n\t"       // + EPS on carry (cannot wrap again: the wrapped high limb is < 2^13)
            "addc.u32 %1, %1, 0;\n\t}"
            : "=&r"(o0), "=&r"(o1) : "r"((u32)t), "r"((u32)(t >> 32)), "r"(b0));
        s[r] = ((u64)o1 << 32) | o0;
    }
}
ZKB_D void mds_layer(u64* s) { mds_layer_rc(s, c_rc2 + 24 * P_ROUNDS); }   // the all-zero tail of the table

// The round constants of round r+1 are folded into the MDS layer of round r (the first round's are added up
// front; c_rc2 ends with 24 zeros so the last round folds nothing). ONE loop over the 30 rounds with a
// warp-uniform full/partial switch, and the 12 S-boxes of a full round done as 3 passes of 4 with a register
// rotation, keep the kernel's code inside the SM's 32 KB instruction cache: with separate unrolled copies per
// round type (64 KB) the leaf kernel spent most of its issue slots stalled on instruction fetch
// (ncu: stall_no_instruction 7.8 per issue, ICC hit rate 61 %).
ZKB_D void poseidon_permute(u64* s) {
#pragma unroll
    for (int i = 0; i < 12; ++i) s[i] = gl_add_lazy_c(s[i], c_rc[i]);
#pragma unroll 1
    for (int r = 0; r < P_ROUNDS; ++r) {
        if (r < P_HALF_FULL || r >= P_HALF_FULL + P_PARTIAL) {
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
#pragma unroll
                for (int i = 0; i < 4; ++i) s[i] = gl_sbox7(s[i]);
                u64 t0 = s[0], t1 = s[1], t2 = s[2], t3 = s[3];
#pragma unroll
                for (int i = 0; i < 8; ++i) s[i] = s[i + 4];
                s[8] = t0; s[9] = t1; s[10] = t2; s[11] = t3;
            }
        } else {
            s[0] = gl_sbox7(s[0]);
        }
        mds_layer_rc(s, c_rc2 + 24 * (r + 1));
    }
}
#endif  // __CUDACC__

}  // namespace zkb
