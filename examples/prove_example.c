/* Minimal C client of libzkb200.so: builds a synthetic wormhole-shaped circuit, proves it once through the C ABI and writes
 * the proof bytes — what a non-Python host (the Rust shim in rust/zkb200, or any C/C++ caller) does.
 *   gcc -std=c11 -O2 -Iinclude examples/prove_example.c -Lzk-circuits_b200 -lzkb200 -lzkb200_synth -Wl,-rpath,$PWD/zk-circuits_b200 -o prove_example
 *   ./prove_example [zk=1] [out.bin]                                                                                   */
#include <stdio.h>
#include <stdlib.h>
#include "zkb200.h"
#include "zkb200_synth.h" /* synthetic workload (libzkb200_synth.so): stands in for the Rust circuit builder + witness generator */

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != ZKB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, zkb_last_error()); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    int zk = argc > 1 ? atoi(argv[1]) : 1;
    if (zkb_device_count() == 0) { fprintf(stderr, "no CUDA device (%s)\n", zkb_version()); return 2; }
    zkb_synth* s = NULL;
    if (zkb_synth_create(0, zk, 6, 5, 6, 3, 5, 42, &s)) { fprintf(stderr, "synth: %s\n", zkb_synth_last_error()); return 1; }           /* tiny row mix: 6 Poseidon, 5 BaseSum, 6 Arithmetic rows */
    size_t n = zkb_synth_degree(s), clen = zkb_synth_common_len(s);
    uint8_t* common = malloc(clen);
    uint64_t* cs = malloc(84 * n * sizeof(uint64_t));
    uint64_t* wires = malloc(135 * n * sizeof(uint64_t));
    uint64_t pis[5];
    if (zkb_synth_get(s, common, cs, wires, pis)) return 1;
    zkb_circuit* c = NULL;
    CHECK(zkb_circuit_create(common, clen, cs, /*is_values=*/1, /*digest=*/NULL, /*device=*/0, &c));
    size_t cap = zkb_proof_size(c), len = 0;
    uint8_t* proof = malloc(cap);
    /* ZKB_SALTS_FROM_SEED makes the run reproducible (the test suite compares the bytes); production callers leave it out and
     * get CSPRNG salts. ZKB_CHECK_WITNESS: refuse to emit a proof for a witness that violates a constraint. */
    CHECK(zkb_prove(c, wires, pis, 5, /*salts=*/NULL, /*salt_seed=*/7, ZKB_POW_MIN | ZKB_SALTS_FROM_SEED | ZKB_CHECK_WITNESS, proof, cap,
                    &len));
    float ms[ZKB_NUM_TIMINGS];
    int nt = zkb_last_timings(c, ms, ZKB_NUM_TIMINGS);
    printf("proof %zu bytes, n = %zu, device total %.3f ms, %llu kernels launched\n", len, n, nt > 12 ? ms[12] : 0.f,
           zkb_kernel_launch_count());
    if (argc > 2) { FILE* f = fopen(argv[2], "wb"); if (!f) return 3; fwrite(proof, 1, len, f); fclose(f); }
    /* error path: a buffer that is too small reports the required size */
    size_t need = 0;
    int rc = zkb_prove(c, wires, pis, 5, NULL, 0, ZKB_POW_MIN, proof, 16, &need);
    if (rc != ZKB_E_BUFFER || need != cap) { fprintf(stderr, "expected ZKB_E_BUFFER with the size, got %d / %zu\n", rc, need); return 4; }
    /* the asynchronous engine: 2 prover contexts, 3 pinned witness slots, 3 proofs in flight from this one thread; the
     * first must equal the blocking call's bytes (same witness, same salt stream) */
    zkb_engine* e = NULL;
    CHECK(zkb_engine_create(common, clen, cs, 1, NULL, 0, /*contexts=*/2, /*slots=*/3, &e));
    uint8_t* eproof[3];
    int slot[3];
    for (int i = 0; i < 3; ++i) {
        uint64_t* buf = NULL;
        slot[i] = zkb_engine_acquire(e, &buf);
        if (slot[i] < 0) { fprintf(stderr, "zkb_engine_acquire -> %d: %s\n", slot[i], zkb_last_error()); return 1; }
        for (size_t k = 0; k < 135 * n; ++k) buf[k] = wires[k];          /* a witness generator would write here directly */
        eproof[i] = malloc(cap);
        CHECK(zkb_engine_submit(e, slot[i], pis, 5, NULL, 7 + (uint64_t)i, ZKB_POW_MIN | ZKB_SALTS_FROM_SEED | ZKB_CHECK_WITNESS,
                                eproof[i], cap));
    }
    for (int i = 0; i < 3; ++i) {
        size_t elen = 0;
        CHECK(zkb_engine_wait(e, slot[i], &elen));
        if (elen != len) { fprintf(stderr, "engine proof %d has %zu bytes, expected %zu\n", i, elen, len); return 5; }
    }
    for (size_t k = 0; k < len; ++k)
        if (eproof[0][k] != proof[k]) { fprintf(stderr, "engine proof differs from zkb_prove at byte %zu\n", k); return 6; }
    printf("engine: 3 proofs, first one byte-identical to zkb_prove\n");
    for (int i = 0; i < 3; ++i) free(eproof[i]);
    CHECK(zkb_engine_destroy(e));
    zkb_circuit_destroy(c);
    zkb_synth_destroy(s);
    free(common); free(cs); free(wires); free(proof);
    return 0;
}
