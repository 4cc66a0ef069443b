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

#define ZKB_MDS_C(i) ((i) == 0 ? 17u : (i) == 1 ? 15u : (i) == 2 ? 41u : (i) == 3 ? 16u : (i) == 4 ? 2u : (i) == 5 ? 28u : \
                      (i) == 6 ? 13u : (i) == 7 ? 13u : (i) == 8 ? 39u : (i) == 9 ? 18u : (i) == 10 ? 34u : 20u)

ZKB_D u64 gl_sbox7(u64 x) {
    u64 x2 = gl_mul_lazy(x, x);
    u64 x3 = gl_mul_lazy(x2, x);
    u64 x4 = gl_mul_lazy(x2, x2);
    return gl_mul_lazy(x3, x4);
}

// out[r] = sum_i s[(i+r)%12] * C[i] + (r==0 ? 8*s[0] : 0); inputs/outputs lazy u64
ZKB_D void mds_layer(u64* s) {
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) { lo[i] = (u32)s[i]; hi[i] = (u32)(s[i] >> 32); }
#pragma unroll
    for (int r = 0; r < 12; ++r) {
        u64 al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            al += (u64)lo[(i + r) % 12] * ZKB_MDS_C(i);
            ah += (u64)hi[(i + r) % 12] * ZKB_MDS_C(i);
        }
        if (r == 0) { al += (u64)lo[0] * 8u; ah += (u64)hi[0] * 8u; }
        // value = al + ah * 2^32  (al, ah < 2^42): fold into (lo64, hi64) then reduce with 2^64 = EPS
        u64 lo64 = al + (ah << 32);
        u64 hi64 = (ah >> 32) + (lo64 < al ? 1u : 0u);     // < 2^11
        u64 t1 = (hi64 << 32) - hi64;
        u64 v = lo64 + t1;
        if (v < t1) v += GL_EPS;
        s[r] = v;
    }
}

ZKB_D void poseidon_permute(u64* s) {
    int rc = 0;
#pragma unroll 1
    for (int r = 0; r < P_HALF_FULL; ++r) {
#pragma unroll
        for (int i = 0; i < 12; ++i) s[i] = gl_sbox7(gl_add_lazy(s[i], c_rc[rc + i]));
        mds_layer(s);
        rc += 12;
    }
#pragma unroll 1
    for (int r = 0; r < P_PARTIAL; ++r) {
#pragma unroll
        for (int i = 1; i < 12; ++i) s[i] = gl_add_lazy(s[i], c_rc[rc + i]);
        s[0] = gl_sbox7(gl_add_lazy(s[0], c_rc[rc]));
        mds_layer(s);
        rc += 12;
    }
#pragma unroll 1
    for (int r = 0; r < P_HALF_FULL; ++r) {
#pragma unroll
        for (int i = 0; i < 12; ++i) s[i] = gl_sbox7(gl_add_lazy(s[i], c_rc[rc + i]));
        mds_layer(s);
        rc += 12;
    }
}
#endif  // __CUDACC__

}  // namespace zkb
