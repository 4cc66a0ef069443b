// Host check of the limb-form Poseidon (csrc/poseidon.cuh: poseidon_permute_limbs, mds_limb12, limb_*) against the
// straightforward u128 permutation of csrc/host_transcript.hpp. The same source is what the CUDA kernels compile.
// Build/run: see tests/test_native_host.py. Exit code 0 = all cases equal.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../zk-circuits_b200/csrc/poseidon.cuh"
#include "../../zk-circuits_b200/csrc/poseidon_consts.hpp"
namespace zkb { }
using namespace zkb;

static u64 sbox_ref(u64 x) { u64 x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x2, x); return gl_mul(x3, x4); }
static void permute_ref(u64* s) {
    static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    const u64* rc = host_round_constants();
    for (int r = 0; r < 30; ++r) {
        for (int i = 0; i < 12; ++i) s[i] = gl_add(s[i], rc[12 * r + i]);
        if (r < 4 || r >= 26) { for (int i = 0; i < 12; ++i) s[i] = sbox_ref(s[i]); } else s[0] = sbox_ref(s[0]);
        u64 out[12];
        for (int o = 0; o < 12; ++o) {
            unsigned __int128 acc = o == 0 ? (unsigned __int128)s[0] * 8 : 0;
            for (int i = 0; i < 12; ++i) acc += (unsigned __int128)s[(i + o) % 12] * C[i];
            out[o] = gl_canon(gl_reduce128_lazy((u64)acc, (u64)(acc >> 64)));
        }
        for (int o = 0; o < 12; ++o) s[o] = out[o];
    }
}
static int check(const u64* in, const char* what) {
    u64 a[12], b[12];
    u32 o0[12], o1[12], o2[12];
    for (int i = 0; i < 12; ++i) { a[i] = in[i]; limb_split(in[i], o0[i], o1[i], o2[i]); }
    permute_ref(a);
    poseidon_permute_limbs(o0, o1, o2, host_round_constant_limbs());
    for (int i = 0; i < 12; ++i) b[i] = gl_canon(limb_to_u64(o0[i], o1[i], o2[i]));
    for (int i = 0; i < 12; ++i)
        if (a[i] != b[i]) { std::printf("MISMATCH (%s) word %d: %016llx vs %016llx\n", what, i, (unsigned long long)a[i], (unsigned long long)b[i]); return 1; }
    return 0;
}
int main(int argc, char** argv) {
    int n = argc > 1 ? std::atoi(argv[1]) : 20000;
    int bad = 0;
    u64 s[12];
    const u64 edge[] = {0, 1, GL_P - 1, GL_P - 2, 0xFFFFFFFFull, 0x100000000ull, 0xFFFFFFFF00000000ull, 0x3FFFFFull, 0x400000ull,
                        0xFFFFFFFFFFFull, 0x100000000000ull, 0xFFFFF00000000000ull, 0x7FFFFFFF80000000ull};
    for (u64 e : edge) { for (int i = 0; i < 12; ++i) s[i] = e; bad += check(s, "edge-all"); }
    for (size_t k = 0; k < sizeof(edge) / 8; ++k) { for (int i = 0; i < 12; ++i) s[i] = edge[(k + i) % (sizeof(edge) / 8)]; bad += check(s, "edge-mixed"); }
    std::mt19937_64 rng(12345);
    for (int t = 0; t < n; ++t) {
        for (int i = 0; i < 12; ++i) { u64 v = rng(); s[i] = v >= GL_P ? v - GL_P : v; }
        if (t % 7 == 0) for (int i = 8; i < 12; ++i) s[i] = 0;
        bad += check(s, "random");
    }
    // chained sponge use: overwrite 8 rate words between permutations, capacity stays in raw limb form
    {
        u64 a[12] = {0};
        u32 o0[12] = {0}, o1[12] = {0}, o2[12] = {0};
        for (int blk = 0; blk < 2000; ++blk) {
            for (int i = 0; i < 8; ++i) { u64 v = rng(); v = v >= GL_P ? v - GL_P : v; a[i] = v; limb_split(v, o0[i], o1[i], o2[i]); }
            permute_ref(a);
            poseidon_permute_limbs(o0, o1, o2, host_round_constant_limbs());
            for (int i = 0; i < 12; ++i)
                if (a[i] != gl_canon(limb_to_u64(o0[i], o1[i], o2[i]))) { std::printf("MISMATCH sponge block %d word %d\n", blk, i); ++bad; break; }
        }
    }
    // the MDS layer alone on extreme limb values (bounds of the normalised form)
    {
        static const int C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
        const int lo = -(1 << 21), hi = (1 << 22) + (1 << 21) - 1;
        for (int t = 0; t < 20000; ++t) {
            int x[12]; u32 y[12];
            for (int i = 0; i < 12; ++i) { int m = rng() % 4; x[i] = m == 0 ? lo : m == 1 ? hi : lo + (int)(rng() % (u64)(hi - lo + 1)); y[i] = (u32)x[i]; }
            static const u32 zeros[36] = {0};
            mds_limb12(y, zeros);
            for (int r = 0; r < 12; ++r) {
                long long acc = r == 0 ? 8ll * x[0] : 0;
                for (int i = 0; i < 12; ++i) acc += (long long)x[(i + r) % 12] * C[i];
                if (acc >= (1ll << 31) - (1 << 22) || acc <= -(1ll << 31) + (1 << 22)) { std::printf("MDS bound violated\n"); ++bad; }
                if ((int)y[r] != (int)acc) { std::printf("MDS mismatch\n"); ++bad; break; }
            }
        }
    }
    std::printf(bad ? "FAILED (%d)\n" : "ok\n", bad);
    return bad ? 1 : 0;
}
