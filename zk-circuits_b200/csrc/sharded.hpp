// Multi-GPU sharding of ONE large commitment / quotient across the GPUs of a node with NCCL over NVLink.
// The reference has no distributed path (single process, rayon: /root/reference/wormhole/aggregator/src/circuits/tree.rs:93-103
// is its only parallelism); this is the mode BASELINE.json's north star asks for: "wire/column polynomials and Merkle leaf ranges
// split by GPU, subtree roots plus quotient chunks exchanged with NCCL over NVLink".
//
// Sharding (G ranks, G | 2^rate_bits, cap_height >= rate_bits). In the reference's leaf order LDE coset j is the contiguous leaf
// block bitrev(j), so rank r owns leaf blocks [r B, (r + 1) B), B = 2^rate_bits / G:
//   commit   columns are ALSO split by rank for the interpolation: a rank uploads and inverse-transforms only its column slice
//            into a window of its own memory that every peer has mapped (CUDA IPC, opened once per communicator and size);
//            after one barrier each rank PULLS the other slices out of their owners' windows with its copy engines over
//            NVLink / NVSwitch, nearest rank first, and starts the LDE of a slice the moment its copy has landed (its own
//            slice at once): a G-stage pipeline of 8 n ncols / G-byte copies behind the transforms. It then extends and
//            hashes its own leaf blocks, builds their Merkle subtrees, and ONE all-gather of the 2^cap_height digests
//            completes the cap (and is the barrier after which a window may be overwritten). ZKB_SHARDED_P2P=0, or a box
//            where the IPC mapping fails, selects the NCCL form: an all-gather of the coefficients in four column chunks.
//            (Measured and rejected: LDE kernels reading the peers' windows directly — 2^20 x 100 on 8 GPUs 6.2 ms against
//            3.4 ms with the NCCL gather: the first transform step's 256-byte strided loads are latency-bound over NVLink.)
//   quotient evaluation is coset-local; the degree-n chunks t_m of t(X) = sum_m X^(m n) t_m(X) satisfy, on coset j with
//            c_j = (g w_N^j)^n:  u_j = sum_m c_j^m t_m  (u_j = the interpolant of t on coset j, one local coset iNTT), i.e. an
//            R x R Vandermonde per coefficient index. ONE all-to-all sends every rank the slice [r n / G, (r + 1) n / G) of all
//            u_j; the rank then solves for its slice of every chunk: t_m[k] = c_0^-m / R sum_j w_R^(-j m) u_j[k].
// NCCL is loaded at run time (dlopen "libnccl.so.2", or $ZKB_NCCL_LIB): libzkb200.so itself has no link-time dependency on it.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <vector>
#include "prover.hpp"

namespace zkb {

constexpr size_t COMM_ID_BYTES = 128;

class Comm {
public:
    static void unique_id(uint8_t out[COMM_ID_BYTES]);
    Comm(const uint8_t id[COMM_ID_BYTES], int nranks, int rank, int device);
    ~Comm();
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    int nranks() const { return nranks_; }
    int rank() const { return rank_; }

    // cap_out: all 2^cap_height digests on every rank. times_ms: {lde (iNTT + gather + LDE), merkle (trees + cap gather),
    // coefficient all-gather alone (overlapped with the LDE)}
    void commit(const u64* values_host, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, int reps, u64* cap_out,
                float* times_ms);
    // q_values_host: this rank's blocks of the quotient evaluations, [nch][B n] in leaf order (block i = leaf block rank B + i);
    // chunks_out_host: [nch][R][n / G] — coefficients [rank n / G, (rank + 1) n / G) of chunk m of challenge ch
    void quotient_chunks(const u64* q_values_host, size_t nch, size_t n, unsigned rate_bits, u64* chunks_out_host, float* times_ms);

    bool peer_windows() const { return p2p_ok_; }

private:
    bool ensure_window(size_t words);   // collective; false: no peer mapping on this box (use the NCCL gather)
    void release_window();              // collective
    void barrier();                     // stream-ordered, through NCCL
    void* comm_ = nullptr;      // ncclComm_t
    int nranks_, rank_, device_;
    cudaStream_t st_ = nullptr, comm_st_ = nullptr;
    u64* win_ = nullptr;                // this rank's coefficient window
    size_t win_words_ = 0;
    std::vector<u64*> peer_win_;        // every rank's window as mapped here ([rank_] = win_)
    bool p2p_ok_ = false, p2p_tried_ = false;
    u64* sync_words_ = nullptr;         // nranks_ words for the barrier / agreement all-gathers
};

}  // namespace zkb
