// The AVX2 form of the Fiat-Shamir permutation (csrc/host_poseidon_avx2.cpp) against the scalar form of host_transcript.hpp on
// random and corner states. Host only: g++ -mavx2 for the AVX2 object, plain flags for this file.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../zk-circuits_b200/csrc/field.cuh"
#include "../../zk-circuits_b200/csrc/poseidon_consts.hpp"
namespace zkb {
const u64* host_round_constants_ptr() { return host_round_constants(); }
void h_poseidon_permute_avx2(u64* s);
// scalar reference: the naive round form (add constants, x^7, dense matrix) with 128-bit arithmetic
static void naive(u64* s) {
    static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    const u64* rc = host_round_constants();
    for (int r = 0; r < 30; ++r) {
        for (int i = 0; i < 12; ++i) s[i] = gl_add(s[i], rc[12 * r + i]);
        for (int i = 0; i < ((r < 4 || r >= 26) ? 12 : 1); ++i) { u64 x = s[i], x2 = gl_mul(x, x), x4 = gl_mul(x2, x2); s[i] = gl_mul(gl_mul(x2, x), x4); }
        u64 o[12];
        for (int q = 0; q < 12; ++q) {
            unsigned __int128 acc = q == 0 ? (unsigned __int128)s[0] * 8 : 0;
            for (int i = 0; i < 12; ++i) acc += (unsigned __int128)s[(i + q) % 12] * C[i];
            o[q] = (u64)(acc % GL_P);
        }
        for (int i = 0; i < 12; ++i) s[i] = o[i];
    }
}
}  // namespace zkb
int main(int argc, char** argv) {
    using namespace zkb;
    if (!__builtin_cpu_supports("avx2")) { std::printf("no avx2: skipped ok\n"); return 0; }
    const int n = argc > 1 ? std::atoi(argv[1]) : 2000;
    std::mt19937_64 rng(7);
    const u64 corner[6] = {0, 1, GL_P - 1, 0xFFFFFFFFull, 0xFFFFFFFF00000000ull, 0x8000000000000000ull};
    for (int t = 0; t < n; ++t) {
        u64 a[12], b[12];
        for (int i = 0; i < 12; ++i) a[i] = b[i] = t < 36 ? corner[(t + i * (t / 6 + 1)) % 6] : gl_canon(rng());
        naive(a);
        h_poseidon_permute_avx2(b);
        for (int i = 0; i < 12; ++i)
            if (a[i] != b[i]) { std::printf("mismatch at state %d word %d\n", t, i); return 1; }
    }
    u64 k[12];
    for (int i = 0; i < 12; ++i) k[i] = i;
    h_poseidon_permute_avx2(k);
    if (k[0] != 0xD64E1E3EFC5B8E9Eull || k[3] != 0x613A4F81E81231D2ull) { std::printf("KAT mismatch\n"); return 2; }
    std::printf("%d states ok\n", n);
    return 0;
}
