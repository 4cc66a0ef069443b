// Host pipeline of the B200 prover: owns all device memory of one circuit context and drives the
// kernels in the order of qp-plonky2 1.1.1 `plonk::prover::prove_with_partition_witness`
// (reached from /root/reference/wormhole/prover/src/lib.rs:234-236 and
// /root/reference/wormhole/aggregator/src/circuits/tree.rs:136; stage list SURVEY.md §3.3 (d)-(l)).
#pragma once
#include <cuda_runtime.h>
#include <memory>
#include <string>
#include <map>
#include <vector>
#include "common_data.hpp"
#include "kernels.h"
#include "host_transcript.hpp"

namespace zkb {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArgError : std::runtime_error { using std::runtime_error::runtime_error; };
struct DigestError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ZetaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct UnsatError : std::runtime_error { using std::runtime_error::runtime_error; };
struct NcclError : std::runtime_error { using std::runtime_error::runtime_error; };

// zkb_prove `flags` (include/zkb200.h): low byte = proof-of-work rule, then option bits
constexpr u32 PF_POW_MASK = 0xffu, PF_SALTS_FROM_SEED = 0x100u, PF_CHECK_WITNESS = 0x200u, PF_WITNESS_RESIDENT = 0x400u;
constexpr u32 PF_KNOWN = PF_POW_MASK | PF_SALTS_FROM_SEED | PF_CHECK_WITNESS | PF_WITNESS_RESIDENT;
struct BufferError : std::runtime_error {
    size_t required;
    BufferError(size_t r) : std::runtime_error("output buffer too small"), required(r) {}
};

void cuda_check(cudaError_t e, const char* what);

class DevBuf {
public:
    DevBuf() = default;
    explicit DevBuf(size_t words) { alloc(words); }
    ~DevBuf() { release(); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p_(o.p_), words_(o.words_), owned_(o.owned_) { o.p_ = nullptr; o.words_ = 0; }
    void alloc(size_t words);
    void view(u64* p, size_t words) { release(); p_ = p; words_ = words; owned_ = false; }   // a slice of someone else's allocation
    void release();
    u64* get() const { return p_; }
    size_t words() const { return words_; }
private:
    u64* p_ = nullptr;
    size_t words_ = 0;
    bool owned_ = true;
};

// one committed polynomial batch on the device (PolynomialBatch): coefficients + LDE (leaf order) + tree
struct BatchDev {
    int ncols = 0, salt = 0;
    DevBuf coeffs;     // [ncols][n]        (may alias an external buffer: see coeff_ptr)
    u64* coeff_ptr = nullptr; size_t coeff_stride = 0;
    DevBuf lde;        // [ncols + salt][N]
    DevBuf digests;    // all levels
    size_t cap_offset = 0;   // digest index of the cap level
};

enum Stage { T_WIRES_INTT = 0, T_WIRES_LDE, T_WIRES_MERKLE, T_PP, T_ZS_COMMIT, T_QUOTIENT, T_QUOTIENT_COMMIT, T_OPENINGS,
             T_FRI_COMBINE, T_FRI_COMMIT, T_POW, T_QUERIES, T_TOTAL, T_HOST_TRANSCRIPT, T_HOST_PERMS, T_COUNT };

class Circuit {
public:
    Circuit(const uint8_t* common, size_t len, const u64* const_sigma, bool is_values, const u64* digest, int device);
    ~Circuit();
    Circuit(const Circuit&) = delete;
    Circuit& operator=(const Circuit&) = delete;

    const CommonData& common() const { return cd_; }
    void verifier_only(u64* cap_out, u64 digest_out[4]) const;
    // throws ArgError / BufferError for anything wrong with a prove call's arguments; touches no device state
    void validate_prove_args(const u64* public_inputs, size_t n_pi, u32 flags, const uint8_t* out, size_t cap) const;
    void upload_witness(const u64* wires_host, bool wait = true);
    // One proof, blocking: begin_proof + (sync, advance) until done.
    size_t prove_resident(const u64* public_inputs, size_t n_pi, const u64* salts, u64 salt_seed, u32 flags, uint8_t* out, size_t cap);
    // The same proof as a RESUMABLE job, for a driver that keeps several contexts in flight from one host thread
    // (engine.hpp): begin_proof() queues the first stage on the stream and returns; whenever ready() says the stream has
    // drained, advance() consumes the stage's results on the host (Fiat-Shamir), queues the next stage and returns true once
    // the proof bytes are complete. An exception from either leaves the context idle and reusable (abort_proof()).
    void begin_proof(const u64* public_inputs, size_t n_pi, const u64* salts, u64 salt_seed, u32 flags, uint8_t* out, size_t cap);
    bool ready();
    bool advance();
    void abort_proof();
    bool busy() const { return job_.stage != ST_IDLE; }
    size_t proof_len() const { return job_.len; }
    void wait_stream() { sync(); }

    void partial_products(const u64* wires_host, const u64* betas, const u64* gammas, u64* out_host);
    void quotient(const u64* wires_host, const u64* zs_pp_host, const u64* pis, size_t n_pi, const u64* betas,
                  const u64* gammas, const u64* alphas, u64* out_host);
    float timings[T_COUNT] = {0};
    int device() const { return device_; }

private:
    void init(const u64* const_sigma, bool is_values, const u64* digest);
    void cleanup();
    void commit_batch(BatchDev& b, unsigned batch_id, u64* cap_host);
    void fill_salts(BatchDev& b, unsigned batch_id);
    void run_partial_products(const u64* betas, const u64* gammas);
    void fill_quotient_params(const u64* pi_hash, const u64* betas, const u64* gammas, const u64* alphas);
    void run_quotient(const u64* pi_hash, const u64* betas, const u64* gammas, const u64* alphas);
    void run_witness_check();
    void queue_flag_readback();
    void check_flags();
    void queue_fri_layer();
    void serialize_proof();
    void sync();

    CommonData cd_;
    int device_ = 0;
    cudaStream_t st_ = nullptr;
    cudaEvent_t sync_ev_ = nullptr;       // blocking-sync event (see Circuit::sync)
    size_t n_ = 0, N_ = 0;
    unsigned lg_n_ = 0, lg_N_ = 0;

    // every device buffer of a context is a slice of ONE allocation: cudaMalloc / cudaFree cost 3-15 ms per call on the
    // B200 boxes (35 separate buffers made zkb_circuit_create 100-500 ms and the destructor 300 ms; one arena: a few ms)
    DevBuf arena_;
    BatchDev cs_, wires_, zs_, quot_;
    DevBuf cs_vals_;             // [num_constants + num_routed][n] values over H, natural order (sigmas: partial products;
                                 // all of them: the witness self-check)
    DevBuf wires_vals_;          // [num_wires][n]
    DevBuf zs_vals_;             // [num_zs_pp][n] values, then coefficients in place (aliased by zs_.coeff_ptr)
    DevBuf q_;                   // [nch][N] quotient values -> coefficients (aliased by quot_.coeff_ptr)
    DevBuf pp_scratch_, k_is_dev_;
    DevBuf zpow_;                // 2 * n
    DevBuf openings_dev_;        // 2 * (all polys + nch)
    DevBuf apow_dev_;            // quotient alpha powers [nch][nterms]
    DevBuf fri_apow_;            // 2 * total columns (SoA)
    DevBuf qparams_dev_;         // QuotientParams
    // small circuits (N <= 2^16): the quotient launches run concurrently, sliced over their gates (kernels.h QuotientFork)
    DevBuf qpart_;               // [slots][nch][N]
    QuotientFork qfork_{};
    bool q_sliced_ = false;
    // FRI
    std::vector<DevBuf> fri_coeffs_;   // per layer (+ final): 2 * m_i (SoA a then b)
    std::vector<DevBuf> fri_values_;   // per layer: 2 * 8 * m_i
    std::vector<DevBuf> fri_digests_;
    std::vector<size_t> fri_cap_off_;
    DevBuf pow_dev_;             // 12 state words + 1 result
    DevBuf flag_dev_;            // bit 0: non-canonical input, bit 1: unsatisfied constraint (witness self-check)
    // A tree's chain of level launches (8 short, strictly dependent kernels on fixed buffers) is captured once into a CUDA
    // graph and replayed: one host call per tree, no launch gaps between 10-50 us kernels.
    struct LevelGraph { cudaGraphExec_t exec = nullptr; size_t cap_offset = 0; unsigned long long kernels = 0; };
    std::map<const u64*, LevelGraph> level_graphs_;
    size_t run_merkle_levels(u64* digests, size_t num_leaves, unsigned cap_height);
    bool witness_loaded_ = false;
    DevBuf query_idx_dev_;       // u32 indices: (1 + layers) * nq, packed in u64 words
    DevBuf query_out_dev_;
    // pinned host staging: one block, regions at fixed word offsets computed (exactly) at construction
    u64* h_stage_ = nullptr;
    size_t h_stage_words_ = 0;
    struct StageOff { size_t caps, open, apow, fin, pow, flag, idx, q, qp; } ho_{};
    size_t q_words_ = 0;
    u64 circuit_digest_[4] = {0};
    std::vector<u64> cs_cap_;
    cudaEvent_t ev_[T_COUNT + 1] = {nullptr};

    // ---- the proof in flight ----
    enum { ST_IDLE = 0, ST_WIRES, ST_ZS, ST_QUOTIENT, ST_OPENINGS, ST_FRI_LAYER, ST_FINAL_POLY, ST_POW, ST_QUERIES };
    struct Job {
        int stage = ST_IDLE;
        std::vector<u64> pis;
        const u64* salt_ptr[3] = {nullptr, nullptr, nullptr};
        u64 salt_seed = 0;
        u32 flags = 0;
        u32 salt_key[8] = {0};
        uint8_t* out = nullptr;
        size_t cap = 0, len = 0;
        Challenger ch;
        u64 pi_hash[4] = {0}, betas[2] = {0}, gammas[2] = {0}, alphas[2] = {0};
        ext2 zeta{}, zeta_next{};
        size_t fri_i = 0, m = 0;
        unsigned lg_m = 0;
        u64 shift = 0, pow_base = 0, pow_witness = 0;
        size_t row_off[4] = {0}, path_off[4] = {0};
        std::vector<size_t> leaf_off, lpath_off;
        std::vector<int> lplen;
    } job_;
};

}  // namespace zkb
