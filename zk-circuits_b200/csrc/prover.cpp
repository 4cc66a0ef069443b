#include "prover.hpp"
#include "host_transcript.hpp"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <thread>
#include <sys/random.h>

namespace zkb {

const u64* host_round_constants_ptr() { return host_round_constants(); }

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw CudaError(std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(x) cuda_check((x), #x)

void DevBuf::alloc(size_t words) {
    release();
    if (words == 0) return;
    CK(cudaMalloc(&p_, words * sizeof(u64)));
    words_ = words;
    owned_ = true;
}
void DevBuf::release() {
    if (p_ && owned_) cudaFree(p_);
    p_ = nullptr;
    words_ = 0;
    owned_ = true;
}

namespace {
struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int d) { CK(cudaGetDevice(&prev)); CK(cudaSetDevice(d)); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};
void check_canonical(const u64* v, size_t n, const char* what) {
    for (size_t i = 0; i < n; ++i)
        if (v[i] >= GL_P) throw ArgError(std::string(what) + ": non-canonical field element at index " + std::to_string(i));
}
}  // namespace

Circuit::Circuit(const uint8_t* common, size_t len, const u64* const_sigma, bool is_values, const u64* digest, int device)
    : cd_(parse_common_data(common, len)), device_(device) {
    if (!const_sigma) throw ArgError("const_sigma is null");
    // host-only checks first: nothing is allocated when they fail
    check_canonical(const_sigma, (size_t)(cd_.num_constants + cd_.num_routed_wires) << cd_.degree_bits, "const_sigma");
    if (cd_.num_challenges > 2) throw UnsupportedError("more than two challenges");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) throw ArgError("bad device index");
    // a constructor that throws does not run the destructor: release streams, events, pinned and device memory here
    try {
        init(const_sigma, is_values, digest);
    } catch (...) {
        cleanup();
        throw;
    }
}

void Circuit::init(const u64* const_sigma, bool is_values, const u64* digest) {
    // ZKB_TRACE=1: wall-clock checkpoints of the context build on stderr (which part of zkb_circuit_create costs what)
    const bool trace = std::getenv("ZKB_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (trace) std::fprintf(stderr, "[zkb create] %-28s %8.3f ms\n", what,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    DeviceGuard g(device_);
    device_tables_init(device_);
    CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    for (auto& e : ev_) CK(cudaEventCreate(&e));
    CK(cudaEventCreateWithFlags(&sync_ev_, cudaEventBlockingSync | cudaEventDisableTiming));
    lg_n_ = (unsigned)cd_.degree_bits;
    lg_N_ = lg_n_ + (unsigned)cd_.rate_bits;
    n_ = size_t(1) << lg_n_;
    N_ = size_t(1) << lg_N_;
    const int ncs = (int)(cd_.num_constants + cd_.num_routed_wires), nw = (int)cd_.num_wires;
    const int nzp = (int)cd_.num_zs_pp(), nq = (int)cd_.num_quotient_polys(), nch = (int)cd_.num_challenges;
    const int salt = (int)cd_.salt_size();
    const unsigned cap_h = (unsigned)cd_.cap_height;
    mark("stream, events, tables");

    // plan every device buffer, then carve them out of one allocation (256-byte aligned slices)
    std::vector<std::pair<DevBuf*, size_t>> plan;
    auto want = [&](DevBuf& b, size_t words) { plan.push_back({&b, words}); };
    auto init_batch = [&](BatchDev& b, int ncols, int s, bool own_coeffs) {
        b.ncols = ncols;
        b.salt = s;
        if (own_coeffs) want(b.coeffs, (size_t)ncols * n_);
        want(b.lde, (size_t)(ncols + s) * N_);
        want(b.digests, merkle_digest_count(N_, cap_h) * 4);
    };
    init_batch(cs_, ncs, 0, true);
    init_batch(wires_, nw, salt, true);
    init_batch(zs_, nzp, salt, false);
    init_batch(quot_, nq, salt, false);
    want(cs_vals_, (size_t)ncs * n_);
    want(wires_vals_, (size_t)nw * n_);
    want(zs_vals_, (size_t)nzp * n_);
    want(q_, (size_t)nch * N_);
    if (lg_N_ <= 16) {           // fewer than ~3 warps per scheduler in a one-thread-per-point launch: slice + overlap
        size_t nrec = 0;
        for (auto& gi : cd_.gates) nrec += gi.is_recursion_gate();
        want(qpart_, ((size_t)nch + 1 + nrec + 1) * nch * N_);
        for (auto& a : qfork_.aux) CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&qfork_.fork, cudaEventDisableTiming));
        for (auto& e : qfork_.join) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        q_sliced_ = true;
    }
    want(pp_scratch_, partial_products_scratch_words((int)cd_.num_routed_wires, (int)cd_.quotient_degree_factor, nch, lg_n_));
    want(k_is_dev_, cd_.k_is.size());
    want(zpow_, 4 * n_);
    const int nall = ncs + nw + nzp + nq;
    want(openings_dev_, 2 * (size_t)(nall + nch));
    const int nterms = nch * (2 + (int)cd_.num_partial_products) + (int)cd_.num_gate_constraints;
    want(apow_dev_, (size_t)nch * nterms);
    want(fri_apow_, 2 * (size_t)nall);
    want(qparams_dev_, (sizeof(QuotientParams) + 7) / 8);
    // FRI layers
    size_t m = n_;
    const size_t L = cd_.reduction_arity_bits.size();
    fri_coeffs_.resize(L + 1);
    fri_values_.resize(L);
    fri_digests_.resize(L);
    for (size_t i = 0; i <= L; ++i) {
        want(fri_coeffs_[i], 2 * m);
        if (i < L) {
            size_t M = m << cd_.rate_bits;
            want(fri_values_[i], 2 * M);
            size_t leaves = M >> cd_.reduction_arity_bits[i];
            want(fri_digests_[i], merkle_digest_count(leaves, cap_h) * 4);
            fri_cap_off_.push_back(0);
            m >>= cd_.reduction_arity_bits[i];
        }
    }
    want(pow_dev_, 16);
    want(flag_dev_, 1);
    const size_t nqr = cd_.num_query_rounds;
    const size_t idx_words = ((1 + L) * nqr + 1) / 2 + 1;
    want(query_idx_dev_, idx_words);
    // query gather buffer
    size_t qwords = 0;
    {
        size_t widths[4] = {(size_t)ncs, (size_t)nw + salt, (size_t)nzp + salt, (size_t)nq + salt};
        for (size_t w : widths) qwords += nqr * (w + 4 * (lg_N_ - cap_h));
        unsigned bits = lg_N_;
        for (u64 ab : cd_.reduction_arity_bits) {
            bits -= (unsigned)ab;
            qwords += nqr * ((size_t(2) << ab) + 4 * (bits >= cap_h ? bits - cap_h : 0));
        }
    }
    q_words_ = qwords;
    want(query_out_dev_, qwords);
    {
        auto round32 = [](size_t w) { return (w + 31) & ~size_t(31); };
        size_t total = 0;
        for (auto& pw : plan) total += round32(pw.second);
        arena_.alloc(total);
        size_t off = 0;
        for (auto& pw : plan) {
            if (pw.second) pw.first->view(arena_.get() + off, pw.second);
            off += round32(pw.second);
        }
    }
    for (BatchDev* b : {&cs_, &wires_}) { b->coeff_ptr = b->coeffs.get(); b->coeff_stride = n_; }
    zs_.coeff_ptr = zs_vals_.get(); zs_.coeff_stride = n_;
    quot_.coeff_ptr = q_.get(); quot_.coeff_stride = n_;     // chunk (ch, m) = q[ch*N + m*n ..]
    if (q_sliced_) qfork_.part = qpart_.get();
    mark("device buffers");
    // pinned staging: every region sized from the circuit's own parameters (cap height, layers, query rounds)
    {
        const size_t cap_words = size_t(4) << cap_h;
        size_t off = 0;
        auto region = [&](size_t words) { size_t at = off; off += (words + 7) & ~size_t(7); return at; };
        ho_.caps = region((3 + L) * cap_words);
        ho_.open = region(2 * (size_t)(nall + nch));
        ho_.apow = region(2 * (size_t)nall);
        ho_.fin = region(2 * cd_.final_poly_len());
        ho_.pow = region(16);
        ho_.flag = region(1);
        ho_.idx = region(idx_words);
        ho_.q = region(qwords);
        ho_.qp = region((sizeof(QuotientParams) + 7) / 8 + (size_t)nch * nterms);
        h_stage_words_ = off;
    }
    CK(cudaMallocHost(&h_stage_, h_stage_words_ * sizeof(u64)));
    mark("pinned staging");

    CK(cudaMemsetAsync(flag_dev_.get(), 0, sizeof(u64), st_));
    CK(cudaMemcpyAsync(k_is_dev_.get(), cd_.k_is.data(), cd_.k_is.size() * 8, cudaMemcpyHostToDevice, st_));
    // ---- constants/sigmas commitment (the part of CircuitBuilder::build the prover needs) ----
    CK(cudaMemcpyAsync(cs_.coeff_ptr, const_sigma, (size_t)ncs * n_ * 8, cudaMemcpyHostToDevice, st_));
    if (is_values) {
        CK(cudaMemcpyAsync(cs_vals_.get(), cs_.coeff_ptr, (size_t)ncs * n_ * 8, cudaMemcpyDeviceToDevice, st_));
        launch_intt_natural(cs_.coeff_ptr, n_, cs_.coeff_ptr, n_, ncs, lg_n_, nullptr, st_);
    } else {
        // values over H: evaluate on <w_n> (rate 0, shift 1) then undo the leaf ordering
        launch_lde(cs_.coeff_ptr, n_, cs_vals_.get(), n_, ncs, lg_n_, 0, 1, st_);
        launch_bitrev_permute(cs_vals_.get(), n_, ncs, lg_n_, st_);
    }
    cs_cap_.resize((size_t(4)) << cap_h);
    commit_batch(cs_, 0, h_stage_ + ho_.caps);
    mark("uploads and launches queued");
    sync();
    mark("constants/sigmas commitment");
    std::memcpy(cs_cap_.data(), h_stage_ + ho_.caps, cs_cap_.size() * 8);
    // circuit_digest = hash_no_pad(cap ‖ hash_pad([]) ‖ [degree_bits])   
    std::vector<u64> parts(cs_cap_);
    u64 ds[4];
    h_hash_pad(nullptr, 0, ds);
    parts.insert(parts.end(), ds, ds + 4);
    parts.push_back(cd_.degree_bits);
    h_hash_no_pad(parts.data(), parts.size(), circuit_digest_);
    if (digest) {
        for (int i = 0; i < 4; ++i)
            if (digest[i] != circuit_digest_[i]) throw DigestError("circuit digest mismatch: constants/sigmas do not match the supplied digest");
    }
}

void Circuit::cleanup() {
    cudaSetDevice(device_);
    if (st_) cudaStreamSynchronize(st_);
    for (auto& kv : level_graphs_) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    level_graphs_.clear();
    for (auto& e : ev_) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& a : qfork_.aux) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); a = nullptr; }
    if (qfork_.fork) { cudaEventDestroy(qfork_.fork); qfork_.fork = nullptr; }
    for (auto& e : qfork_.join) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (sync_ev_) { cudaEventDestroy(sync_ev_); sync_ev_ = nullptr; }
    if (h_stage_) { cudaFreeHost(h_stage_); h_stage_ = nullptr; }
    if (st_) { cudaStreamDestroy(st_); st_ = nullptr; }
    arena_.release();
}

Circuit::~Circuit() { cleanup(); }

// Host waits on the stream ~10 times per proof. Spinning (the runtime's default) gives the lowest latency. With more
// proofs in flight in a process than host cores available to it (8 ranks x 8 streams on a 32-core box) the wait polls
// every ~20 us and sleeps in between instead, so that the threads doing Fiat-Shamir work are not starved by spinners:
// measured at 8 GPUs / 32 cores, 8 streams per GPU sleep-polling 1622 proofs/s (98 % of 8 x one GPU), 4 streams per GPU
// spinning 1419, 8 streams yielding 1410 (profiles/r01_bench_n8_sync_modes.json). A blocking (interrupt) wait is much
// worse (708). ZKB_SYNC=spin|yield|block|sleep overrides the choice.
static std::atomic<int> g_proofs_in_flight{0};
static int sync_mode() {      // 0 spin, 1 yield, 2 block, 3 sleep-poll
    if (const char* e = std::getenv("ZKB_SYNC")) {
        if (!std::strcmp(e, "sleep")) return 3;
        if (!std::strcmp(e, "block")) return 2;
        if (!std::strcmp(e, "yield")) return 1;
        if (!std::strcmp(e, "spin")) return 0;
    }
    unsigned cores = std::thread::hardware_concurrency();
    int local_world = 1;
    if (const char* w = std::getenv("LOCAL_WORLD_SIZE")) local_world = std::atoi(w) > 0 ? std::atoi(w) : 1;
    const unsigned share = cores / (unsigned)local_world;
    return (share != 0 && (unsigned)g_proofs_in_flight.load() >= share) ? 3 : 0;
}
void Circuit::sync() {
    const int mode = sync_mode();
    if (mode == 2) {
        CK(cudaEventRecord(sync_ev_, st_));
        CK(cudaEventSynchronize(sync_ev_));
    } else if (mode == 1) {
        cudaError_t e;
        while ((e = cudaStreamQuery(st_)) == cudaErrorNotReady) std::this_thread::yield();
        CK(e);
    } else if (mode == 3) {      // poll every ~20 us and sleep in between: frees the core, costs tens of microseconds per wait
        cudaError_t e;
        while ((e = cudaStreamQuery(st_)) == cudaErrorNotReady) std::this_thread::sleep_for(std::chrono::microseconds(20));
        CK(e);
    } else {
        CK(cudaStreamSynchronize(st_));
    }
}

size_t Circuit::run_merkle_levels(u64* digests, size_t num_leaves, unsigned cap_height) {
    LevelGraph& g = level_graphs_[digests];
    if (!g.exec) {
        // one capture at a time per process: it happens once per tree and context, keeps the launch count of the captured
        // chain exact when several contexts warm up together, and profilers (ncu) crash on concurrent stream captures
        static std::mutex capture_mu;
        std::lock_guard<std::mutex> lk(capture_mu);
        const unsigned long long before = kernel_launch_count();
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
        g.cap_offset = launch_merkle_levels(digests, num_leaves, cap_height, st_);
        CK(cudaStreamEndCapture(st_, &graph));
        g.kernels = kernel_launch_count() - before;      // counted once at capture; every replay adds them again
        if (!graph) return g.cap_offset;                 // a tree with no level above the leaves
        CK(cudaGraphInstantiate(&g.exec, graph, 0));
        CK(cudaGraphDestroy(graph));
    } else {
        kernel_launch_count_add(g.kernels);
    }
    CK(cudaGraphLaunch(g.exec, st_));
    return g.cap_offset;
}

void Circuit::verifier_only(u64* cap_out, u64 digest_out[4]) const {
    if (cap_out) std::memcpy(cap_out, cs_cap_.data(), cs_cap_.size() * 8);
    if (digest_out) std::memcpy(digest_out, circuit_digest_, 32);
}

// salt columns of a blinded batch: the caller's (explicit), the documented seeded stream (ZKB_SALTS_FROM_SEED: parity tests),
// or — the default — ChaCha20 keyed from the OS RNG for this proof
void Circuit::fill_salts(BatchDev& b, unsigned batch_id) {
    if (!b.salt) return;
    u64* sp = b.lde.get() + (size_t)b.ncols * N_;
    const size_t words = (size_t)b.salt * N_;
    if (job_.salt_ptr[batch_id]) {
        CK(cudaMemcpyAsync(sp, job_.salt_ptr[batch_id], words * 8, cudaMemcpyHostToDevice, st_));
        launch_canonical_check(sp, words, reinterpret_cast<unsigned*>(flag_dev_.get()), st_);
    } else if (job_.flags & PF_SALTS_FROM_SEED) {
        launch_salt_fill(sp, N_, N_, job_.salt_seed, batch_id, st_);
    } else {
        launch_salt_fill_csprng(sp, words, job_.salt_key, batch_id, st_);
    }
}

// LDE (+ salts) + Merkle for a batch whose coefficients are in place; cap is copied to cap_host (pinned) asynchronously
void Circuit::commit_batch(BatchDev& b, unsigned batch_id, u64* cap_host) {
    const unsigned cap_h = (unsigned)cd_.cap_height;
    launch_lde(b.coeff_ptr, b.coeff_stride, b.lde.get(), N_, b.ncols, lg_n_, (unsigned)cd_.rate_bits, GL_GEN, st_);
    if (&b != &cs_) fill_salts(b, batch_id);
    launch_merkle_leaves(b.lde.get(), N_, b.ncols + b.salt, N_, b.digests.get(), st_);
    b.cap_offset = run_merkle_levels(b.digests.get(), N_, cap_h);
    CK(cudaMemcpyAsync(cap_host, b.digests.get() + b.cap_offset * 4, (size_t(32)) << cap_h, cudaMemcpyDeviceToHost, st_));
}

// H2D of the wire matrix + canonical check on the device. With wait = false nothing synchronises: the copy and the check
// are queued ahead of the proof's kernels and the flag is read at the proof's first stage boundary (zkb_prove path).
void Circuit::upload_witness(const u64* wires_host, bool wait) {
    if (!wires_host) throw ArgError("wires is null");
    if (busy()) throw ArgError("a proof is in flight on this context");
    DeviceGuard g(device_);
    unsigned* flag = reinterpret_cast<unsigned*>(flag_dev_.get());
    CK(cudaMemsetAsync(flag, 0, sizeof(u64), st_));
    CK(cudaMemcpyAsync(wires_vals_.get(), wires_host, cd_.num_wires * n_ * 8, cudaMemcpyHostToDevice, st_));
    launch_canonical_check(wires_vals_.get(), cd_.num_wires * n_, flag, st_);
    witness_loaded_ = true;
    if (wait) {
        queue_flag_readback();
        sync();
        try {
            check_flags();
        } catch (...) {
            witness_loaded_ = false;
            throw;
        }
    }
}
void Circuit::queue_flag_readback() {
    CK(cudaMemcpyAsync(h_stage_ + ho_.flag, flag_dev_.get(), sizeof(u64), cudaMemcpyDeviceToHost, st_));
}
void Circuit::check_flags() {     // stream must be idle
    const u64 f = h_stage_[ho_.flag];
    if (f & 1) throw ArgError("non-canonical field element in the wires or salts");
    if (f & 2) throw UnsatError("the witness does not satisfy the circuit (a gate or copy constraint fails on the subgroup)");
}

void Circuit::run_partial_products(const u64* betas, const u64* gammas) {
    u64 bg[8];
    const int nch = (int)cd_.num_challenges;
    for (int c = 0; c < nch; ++c) { bg[c] = betas[c]; bg[nch + c] = gammas[c]; }
    launch_partial_products(wires_vals_.get(), n_, cs_vals_.get() + cd_.num_constants * n_, n_, k_is_dev_.get(),
                            (int)cd_.num_routed_wires, (int)cd_.quotient_degree_factor, nch, bg, lg_n_, zs_vals_.get(), n_,
                            pp_scratch_.get(), st_);
}

// QuotientParams + powers of the combination challenges into the pinned staging block, then to the device (stream order)
void Circuit::fill_quotient_params(const u64* pi_hash, const u64* betas, const u64* gammas, const u64* alphas) {
    const int nch = (int)cd_.num_challenges;
    const int nterms = nch * (2 + (int)cd_.num_partial_products) + (int)cd_.num_gate_constraints;
    QuotientParams* qp = reinterpret_cast<QuotientParams*>(h_stage_ + ho_.qp);
    std::memset(qp, 0, sizeof(QuotientParams));
    qp->lg_n = lg_n_; qp->rate_bits = (unsigned)cd_.rate_bits;
    qp->num_wires = (int)cd_.num_wires; qp->num_routed = (int)cd_.num_routed_wires; qp->num_constants = (int)cd_.num_constants;
    qp->num_selectors = (int)cd_.groups.size(); qp->num_challenges = nch; qp->num_partial_products = (int)cd_.num_partial_products;
    qp->qdf = (int)cd_.quotient_degree_factor; qp->num_gates = (int)cd_.gates.size(); qp->num_gate_constraints = (int)cd_.num_gate_constraints;
    for (size_t g = 0; g < cd_.gates.size(); ++g) {
        u64 sel = cd_.selector_indices[g];
        const GateInfo& gi = cd_.gates[g];
        qp->gates[g] = GateDesc{gi.tag, (u32)gi.param, (u32)sel, (u32)cd_.groups[sel].first, (u32)cd_.groups[sel].second, (u32)g,
                                (u32)gi.p2, (u32)gi.p3};
        if (gi.is_recursion_gate()) qp->has_recursion_gates = 1;
        if (gi.tag == GT_COSET_INTERP) {      // one parameter set per circuit (checked at create)
            const u64 w = gl_root_of_unity((unsigned)gi.param);
            u64 x = 1;
            for (size_t k = 0; k < gi.weights.size(); ++k) { qp->bary_w[k] = gi.weights[k]; qp->bary_x[k] = x; x = gl_mul(x, w); }
        }
    }
    for (size_t j = 0; j < cd_.k_is.size(); ++j) qp->k_is[j] = cd_.k_is[j];
    for (int c = 0; c < nch; ++c) { qp->betas[c] = betas[c]; qp->gammas[c] = gammas[c]; qp->alphas[c] = alphas[c]; }
    for (int i = 0; i < 4; ++i) qp->pi_hash[i] = pi_hash[i];
    const unsigned rate = 1u << cd_.rate_bits;
    u64 g_n = gl_pow(GL_GEN, n_), w_rate = gl_root_of_unity((unsigned)cd_.rate_bits);
    for (unsigned j = 0; j < rate; ++j) {
        qp->zh[j] = gl_sub(gl_mul(g_n, gl_pow(w_rate, j)), 1);
        qp->zh_inv[j] = gl_inv(qp->zh[j]);
    }
    u64* ap = h_stage_ + ho_.qp + (sizeof(QuotientParams) + 7) / 8;
    for (int c = 0; c < nch; ++c) {
        u64 p = 1;
        for (int k = 0; k < nterms; ++k) { ap[(size_t)c * nterms + k] = p; p = gl_mul(p, alphas[c]); }
    }
    CK(cudaMemcpyAsync(qparams_dev_.get(), qp, sizeof(QuotientParams), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(apow_dev_.get(), ap, (size_t)nch * nterms * 8, cudaMemcpyHostToDevice, st_));
}

void Circuit::run_quotient(const u64* pi_hash, const u64* betas, const u64* gammas, const u64* alphas) {
    const int nch = (int)cd_.num_challenges;
    const int nterms = nch * (2 + (int)cd_.num_partial_products) + (int)cd_.num_gate_constraints;
    fill_quotient_params(pi_hash, betas, gammas, alphas);
    const QuotientParams* qp = reinterpret_cast<const QuotientParams*>(h_stage_ + ho_.qp);
    launch_quotient(reinterpret_cast<const QuotientParams*>(qparams_dev_.get()), *qp, apow_dev_.get(), nterms, cs_.lde.get(), N_,
                    wires_.lde.get(), N_, zs_.lde.get(), N_, q_.get(), N_, st_, q_sliced_ ? &qfork_ : nullptr);
    // values on the coset (leaf order) -> coefficients; chunks of n are the committed quotient polynomials
    launch_coset_intt_bitrev(q_.get(), N_, nch, lg_N_, GL_GEN, st_);
}

// ZKB_CHECK_WITNESS (SURVEY §8b's optional self-check, done where the data already is): every constraint of the vanishing
// polynomial — gates with their selector filters, Z(1) = 1, the partial-product chain and its wrap-around, i.e. the copy
// constraints — evaluated at the n points of H from the VALUES of constants/sigmas, wires and Z/partial products, combined
// with powers of a check challenge drawn after the wires are committed. A proof can verify iff all of them vanish on H, so a
// non-zero sum anywhere means the reference verifier would reject: the call returns ZKB_E_UNSAT instead of proof bytes
// (the CPU prover does not notice either — SURVEY §8 a8 — its callers get `Err` from witness generation).
void Circuit::run_witness_check() {
    const int nch = (int)cd_.num_challenges;
    const int nterms = nch * (2 + (int)cd_.num_partial_products) + (int)cd_.num_gate_constraints;
    u64 seed_in[6] = {job_.betas[0], job_.gammas[0], job_.betas[nch - 1], job_.gammas[nch - 1], 0x5a4b4232u, 0}, chk[4];
    h_hash_no_pad(seed_in, 6, chk);
    fill_quotient_params(job_.pi_hash, job_.betas, job_.gammas, chk);
    const QuotientParams* qp = reinterpret_cast<const QuotientParams*>(h_stage_ + ho_.qp);
    launch_constraint_check(reinterpret_cast<const QuotientParams*>(qparams_dev_.get()), *qp, apow_dev_.get(), nterms,
                            cs_vals_.get(), n_, wires_vals_.get(), n_, zs_vals_.get(), n_, q_.get(), N_,
                            reinterpret_cast<unsigned*>(flag_dev_.get()), st_);
}

void Circuit::partial_products(const u64* wires_host, const u64* betas, const u64* gammas, u64* out_host) {
    if (!wires_host || !betas || !gammas || !out_host) throw ArgError("null argument");
    if (busy()) throw ArgError("a proof is in flight on this context");
    DeviceGuard g(device_);
    CK(cudaMemcpyAsync(wires_vals_.get(), wires_host, cd_.num_wires * n_ * 8, cudaMemcpyHostToDevice, st_));
    witness_loaded_ = false;
    run_partial_products(betas, gammas);
    CK(cudaMemcpyAsync(out_host, zs_vals_.get(), cd_.num_zs_pp() * n_ * 8, cudaMemcpyDeviceToHost, st_));
    sync();
}

void Circuit::quotient(const u64* wires_host, const u64* zs_pp_host, const u64* pis, size_t n_pi, const u64* betas,
                       const u64* gammas, const u64* alphas, u64* out_host) {
    if (!wires_host || !zs_pp_host || !betas || !gammas || !alphas || !out_host) throw ArgError("null argument");
    if (busy()) throw ArgError("a proof is in flight on this context");
    DeviceGuard g(device_);
    const int nw = (int)cd_.num_wires, nzp = (int)cd_.num_zs_pp();
    CK(cudaMemcpyAsync(wires_vals_.get(), wires_host, (size_t)nw * n_ * 8, cudaMemcpyHostToDevice, st_));
    witness_loaded_ = false;
    CK(cudaMemcpyAsync(zs_vals_.get(), zs_pp_host, (size_t)nzp * n_ * 8, cudaMemcpyHostToDevice, st_));
    launch_intt_natural(wires_vals_.get(), n_, wires_.coeff_ptr, n_, nw, lg_n_, nullptr, st_);
    launch_lde(wires_.coeff_ptr, n_, wires_.lde.get(), N_, nw, lg_n_, (unsigned)cd_.rate_bits, GL_GEN, st_);
    launch_intt_natural(zs_vals_.get(), n_, zs_vals_.get(), n_, nzp, lg_n_, nullptr, st_);
    launch_lde(zs_vals_.get(), n_, zs_.lde.get(), N_, nzp, lg_n_, (unsigned)cd_.rate_bits, GL_GEN, st_);
    u64 pih[4];
    h_hash_no_pad(pis, n_pi, pih);
    run_quotient(pih, betas, gammas, alphas);
    CK(cudaMemcpy2DAsync(out_host, N_ * 8, q_.get(), N_ * 8, N_ * 8, cd_.num_challenges, cudaMemcpyDeviceToHost, st_));
    sync();
}

namespace {
struct ByteWriter {
    uint8_t* p;
    size_t pos = 0;
    void w64(u64 v) { std::memcpy(p + pos, &v, 8); pos += 8; }   // little-endian host
    void words(const u64* v, size_t n) { std::memcpy(p + pos, v, n * 8); pos += n * 8; }
    void byte(uint8_t b) { p[pos++] = b; }
};
void os_random(void* buf, size_t len) {
    uint8_t* p = static_cast<uint8_t*>(buf);
    while (len) {
        ssize_t got = getrandom(p, len, 0);
        if (got < 0) {
            if (errno == EINTR) continue;
            throw CudaError("getrandom failed: cannot key the salt generator");
        }
        p += got;
        len -= (size_t)got;
    }
}
}  // namespace

void Circuit::validate_prove_args(const u64* public_inputs, size_t n_pi, u32 flags, const uint8_t* out, size_t cap) const {
    if (n_pi != cd_.num_public_inputs) throw ArgError("wrong number of public inputs");
    if (n_pi && !public_inputs) throw ArgError("public_inputs is null");
    if ((flags & PF_POW_MASK) != 0) throw ArgError("unknown pow_rule");
    if (flags & ~PF_KNOWN) throw ArgError("unknown flag bits");
    const size_t psize = cd_.proof_size();
    if (!out || cap < psize) throw BufferError(psize);
    check_canonical(public_inputs, n_pi, "public_inputs");
}

size_t Circuit::prove_resident(const u64* public_inputs, size_t n_pi, const u64* salts, u64 salt_seed, u32 flags,
                               uint8_t* out, size_t cap) {
    try {
        begin_proof(public_inputs, n_pi, salts, salt_seed, flags, out, cap);
        for (;;) {
            sync();
            if (advance()) break;
        }
    } catch (...) {
        abort_proof();
        throw;
    }
    return job_.len;
}

bool Circuit::ready() {
    cudaError_t e = cudaStreamQuery(st_);
    if (e == cudaErrorNotReady) return false;
    CK(e);
    return true;
}
void Circuit::abort_proof() {
    if (st_) {
        cudaSetDevice(device_);
        cudaStreamSynchronize(st_);
        for (auto& a : qfork_.aux) if (a) cudaStreamSynchronize(a);
    }
    if (job_.stage != ST_IDLE) --g_proofs_in_flight;
    job_.stage = ST_IDLE;
}

// Stage 0 of a proof: transcript start, wires commitment queued.
void Circuit::begin_proof(const u64* public_inputs, size_t n_pi, const u64* salts, u64 salt_seed, u32 flags, uint8_t* out,
                          size_t cap) {
    if (busy()) throw ArgError("a proof is in flight on this context");
    validate_prove_args(public_inputs, n_pi, flags, out, cap);
    if (!witness_loaded_) throw ArgError("no witness on the device: call zkb_witness_upload first (or zkb_prove)");
    DeviceGuard g(device_);
    Job& j = job_;
    j = Job();
    j.pis.assign(public_inputs, public_inputs + n_pi);
    j.salt_seed = salt_seed; j.flags = flags; j.out = out; j.cap = cap;
    if (cd_.salt_size()) {
        if (salts) for (int b = 0; b < 3; ++b) j.salt_ptr[b] = salts + (size_t)b * 4 * N_;
        else if (!(flags & PF_SALTS_FROM_SEED)) os_random(j.salt_key, sizeof(j.salt_key));
    }
    h_hash_no_pad(j.pis.data(), n_pi, j.pi_hash);
    j.ch.observe_many(circuit_digest_, 4);
    j.ch.observe_many(j.pi_hash, 4);

    const int nw = wires_.ncols;
    const unsigned cap_h = (unsigned)cd_.cap_height;
    const size_t cap_words = size_t(4) << cap_h;
    CK(cudaEventRecord(ev_[0], st_));
    launch_intt_natural(wires_vals_.get(), n_, wires_.coeff_ptr, n_, nw, lg_n_, nullptr, st_);
    CK(cudaEventRecord(ev_[T_WIRES_INTT + 1], st_));
    launch_lde(wires_.coeff_ptr, n_, wires_.lde.get(), N_, nw, lg_n_, (unsigned)cd_.rate_bits, GL_GEN, st_);
    CK(cudaEventRecord(ev_[T_WIRES_LDE + 1], st_));      // exactly the coset-LDE launch: bench.py's roofline kernel
    fill_salts(wires_, 0);
    launch_merkle_leaves(wires_.lde.get(), N_, nw + wires_.salt, N_, wires_.digests.get(), st_);
    wires_.cap_offset = run_merkle_levels(wires_.digests.get(), N_, cap_h);
    CK(cudaMemcpyAsync(h_stage_ + ho_.caps, wires_.digests.get() + wires_.cap_offset * 4, cap_words * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaEventRecord(ev_[T_WIRES_MERKLE + 1], st_));
    queue_flag_readback();
    ++g_proofs_in_flight;
    j.stage = ST_WIRES;
}

// one FRI commit-phase layer: LDE of the current coefficients on the shifted coset, 2^arity_bits values per leaf, tree, cap
void Circuit::queue_fri_layer() {
    Job& j = job_;
    const size_t i = j.fri_i;
    const unsigned ab = (unsigned)cd_.reduction_arity_bits[i], cap_h = (unsigned)cd_.cap_height, rate_bits = (unsigned)cd_.rate_bits;
    const size_t M = j.m << rate_bits, cap_words = size_t(4) << cap_h;
    u64* ca = fri_coeffs_[i].get();
    u64* va = fri_values_[i].get();
    launch_lde(ca, j.m, va, M, 2, j.lg_m, rate_bits, j.shift, st_);
    const size_t leaves = M >> ab;
    launch_merkle_leaves_ext(va, va + M, 1 << ab, leaves, fri_digests_[i].get(), st_);
    fri_cap_off_[i] = run_merkle_levels(fri_digests_[i].get(), leaves, cap_h);
    CK(cudaMemcpyAsync(h_stage_ + ho_.caps + (3 + i) * cap_words, fri_digests_[i].get() + fri_cap_off_[i] * 4, cap_words * 8,
                       cudaMemcpyDeviceToHost, st_));
}

// Called with the stream idle: the stage's results are in the pinned block. Absorb them into the transcript, draw the
// next challenges, queue the next stage. Returns true when the proof is complete.
bool Circuit::advance() {
    Job& j = job_;
    if (j.stage == ST_IDLE) throw ArgError("no proof in flight");
    DeviceGuard g(device_);
    const int ncs = cs_.ncols, nw = wires_.ncols, nzp = zs_.ncols, nq = quot_.ncols, nch = (int)cd_.num_challenges;
    const int nall = ncs + nw + nzp + nq;
    const unsigned cap_h = (unsigned)cd_.cap_height;
    const size_t cap_words = size_t(4) << cap_h;
    const size_t L = cd_.reduction_arity_bits.size();
    const int nqr = (int)cd_.num_query_rounds;
    u64* h_caps = h_stage_ + ho_.caps;
    u64* h_open = h_stage_ + ho_.open;
    u64* h_final = h_stage_ + ho_.fin;
    u64* h_pow = h_stage_ + ho_.pow;
    check_flags();

    switch (j.stage) {
    case ST_WIRES: {
        j.ch.observe_many(h_caps, cap_words);
        for (int c = 0; c < nch; ++c) j.betas[c] = j.ch.get();
        for (int c = 0; c < nch; ++c) j.gammas[c] = j.ch.get();
        // (f) partial products and Z, (g) commit
        run_partial_products(j.betas, j.gammas);
        CK(cudaEventRecord(ev_[T_PP + 1], st_));
        if (j.flags & PF_CHECK_WITNESS) run_witness_check();
        launch_intt_natural(zs_vals_.get(), n_, zs_vals_.get(), n_, nzp, lg_n_, nullptr, st_);
        commit_batch(zs_, 1, h_caps + cap_words);
        CK(cudaEventRecord(ev_[T_ZS_COMMIT + 1], st_));
        queue_flag_readback();
        j.stage = ST_ZS;
        return false;
    }
    case ST_ZS: {
        j.ch.observe_many(h_caps + cap_words, cap_words);
        for (int c = 0; c < nch; ++c) j.alphas[c] = j.ch.get();
        // (h) quotient, (i) commit
        run_quotient(j.pi_hash, j.betas, j.gammas, j.alphas);
        CK(cudaEventRecord(ev_[T_QUOTIENT + 1], st_));
        commit_batch(quot_, 2, h_caps + 2 * cap_words);
        CK(cudaEventRecord(ev_[T_QUOTIENT_COMMIT + 1], st_));
        queue_flag_readback();
        j.stage = ST_QUOTIENT;
        return false;
    }
    case ST_QUOTIENT: {
        j.ch.observe_many(h_caps + 2 * cap_words, cap_words);
        j.zeta = j.ch.get_ext();
        if (e_eq(e_pow2k(j.zeta, lg_n_), e_from(1))) throw ZetaError("Opening point is in the subgroup.");
        j.zeta_next = e_mul_base(j.zeta, gl_root_of_unity(lg_n_));
        // (j) openings
        OpeningsArgs oa;
        oa.seg[0] = {cs_.coeff_ptr, n_, ncs, 0};
        oa.seg[1] = {wires_.coeff_ptr, n_, nw, 0};
        oa.seg[2] = {zs_.coeff_ptr, n_, nzp, 0};
        oa.seg[3] = {quot_.coeff_ptr, n_, nq, 0};
        oa.seg[4] = {zs_.coeff_ptr, n_, nch, 1};          // Z polynomials again, at g * zeta
        oa.nseg = 5; oa.lg_n = lg_n_; oa.pw = zpow_.get(); oa.out = openings_dev_.get();
        launch_openings(oa, j.zeta, j.zeta_next, st_);
        CK(cudaMemcpyAsync(h_open, openings_dev_.get(), 2 * (size_t)(nall + nch) * 8, cudaMemcpyDeviceToHost, st_));
        CK(cudaEventRecord(ev_[T_OPENINGS + 1], st_));
        j.stage = ST_OPENINGS;
        return false;
    }
    case ST_OPENINGS: {
        // transcript order: constants, sigmas, wires, zs, partial products, quotient, then zs_next — which is the
        // device order (cs, wires, zs_pp, quot, zs_next)
        j.ch.observe_many(h_open, 2 * (size_t)(nall + nch));
        // (k) prove_openings: batch combination -> final polynomial
        const ext2 alpha = j.ch.get_ext();
        u64* h_apow = h_stage_ + ho_.apow;                       // SoA: a[nall], b[nall]
        ext2 reduced0 = e_make(0, 0), reduced1 = e_make(0, 0);
        ext2 p = e_from(1);
        for (int k = 0; k < nall; ++k) {
            h_apow[k] = p.a;
            h_apow[nall + k] = p.b;
            reduced0 = e_add(reduced0, e_mul(p, e_make(h_open[2 * k], h_open[2 * k + 1])));
            if (k < nch) reduced1 = e_add(reduced1, e_mul(p, e_make(h_open[2 * (nall + k)], h_open[2 * (nall + k) + 1])));
            p = e_mul(p, alpha);
        }
        CK(cudaMemcpyAsync(fri_apow_.get(), h_apow, 2 * (size_t)nall * 8, cudaMemcpyHostToDevice, st_));
        FriCombineParams fp;
        fp.lde[0] = cs_.lde.get(); fp.lde[1] = wires_.lde.get(); fp.lde[2] = zs_.lde.get(); fp.lde[3] = quot_.lde.get();
        for (int t = 0; t < 4; ++t) fp.stride[t] = N_;
        fp.ncols[0] = ncs; fp.ncols[1] = nw; fp.ncols[2] = nzp; fp.ncols[3] = nq;
        fp.num_zs = nch;
        fp.alpha = alpha; fp.zeta = j.zeta; fp.zeta_next = j.zeta_next; fp.reduced0 = reduced0; fp.reduced1 = reduced1;
        fp.lg_n = lg_n_;
        u64* fc = fri_coeffs_[0].get();
        launch_fri_combine(fp, fri_apow_.get(), fri_apow_.get() + nall, fc, fc + n_, st_);
        launch_coset_intt_bitrev(fc, n_, 2, lg_n_, GL_GEN, st_);
        CK(cudaEventRecord(ev_[T_FRI_COMBINE + 1], st_));
        // FRI commit phase
        j.shift = GL_GEN; j.m = n_; j.lg_m = lg_n_; j.fri_i = 0;
        if (L > 0) {
            queue_fri_layer();
            j.stage = ST_FRI_LAYER;
        } else {
            CK(cudaMemcpyAsync(h_final, fri_coeffs_[0].get(), 2 * j.m * 8, cudaMemcpyDeviceToHost, st_));
            CK(cudaEventRecord(ev_[T_FRI_COMMIT + 1], st_));
            j.stage = ST_FINAL_POLY;
        }
        return false;
    }
    case ST_FRI_LAYER: {
        const size_t i = j.fri_i;
        const unsigned ab = (unsigned)cd_.reduction_arity_bits[i];
        j.ch.observe_many(h_caps + (3 + i) * cap_words, cap_words);
        const ext2 beta = j.ch.get_ext();
        u64* ca = fri_coeffs_[i].get();
        u64* na = fri_coeffs_[i + 1].get();
        launch_fri_fold(ca, ca + j.m, na, na + (j.m >> ab), j.m >> ab, 1 << ab, beta, st_);
        j.m >>= ab;
        j.lg_m -= ab;
        j.shift = gl_pow(j.shift, u64(1) << ab);
        if (++j.fri_i < L) {
            queue_fri_layer();
        } else {
            CK(cudaMemcpyAsync(h_final, fri_coeffs_[L].get(), 2 * j.m * 8, cudaMemcpyDeviceToHost, st_));
            CK(cudaEventRecord(ev_[T_FRI_COMMIT + 1], st_));
            j.stage = ST_FINAL_POLY;
        }
        return false;
    }
    case ST_FINAL_POLY: {
        const size_t fin = j.m;   // == final_poly_len
        for (size_t k = 0; k < fin; ++k) j.ch.observe_ext(e_make(h_final[k], h_final[fin + k]));
        // proof of work (MIN rule): smallest witness whose response has >= pow_bits leading zeros
        for (int i = 0; i < 12; ++i) h_pow[i] = j.ch.sponge[i];
        for (int i = 0; i < j.ch.in_len; ++i) h_pow[i] = j.ch.in_buf[i];
        h_pow[12] = ~u64(0);
        CK(cudaMemcpyAsync(pow_dev_.get(), h_pow, 13 * 8, cudaMemcpyHostToDevice, st_));
        j.pow_base = 0;
        launch_pow_search(pow_dev_.get(), j.ch.in_len, j.pow_base, u64(1) << 32, cd_.proof_of_work_bits,
                          reinterpret_cast<unsigned long long*>(pow_dev_.get() + 12), st_);
        CK(cudaMemcpyAsync(h_pow + 13, pow_dev_.get() + 12, 8, cudaMemcpyDeviceToHost, st_));
        j.stage = ST_POW;
        return false;
    }
    case ST_POW: {
        if (h_pow[13] == ~u64(0)) {      // nothing in this 2^32 range (probability ~e^-65536 at 16 bits): scan the next one
            const u64 batch = u64(1) << 32;
            j.pow_base += batch;
            if (j.pow_base >= GL_P - batch) throw CudaError("proof of work search exhausted");
            launch_pow_search(pow_dev_.get(), j.ch.in_len, j.pow_base, batch, cd_.proof_of_work_bits,
                              reinterpret_cast<unsigned long long*>(pow_dev_.get() + 12), st_);
            CK(cudaMemcpyAsync(h_pow + 13, pow_dev_.get() + 12, 8, cudaMemcpyDeviceToHost, st_));
            return false;
        }
        j.pow_witness = h_pow[13];
        j.ch.observe(j.pow_witness);
        const u64 resp = j.ch.get();
        if (cd_.proof_of_work_bits && (resp >> (64 - cd_.proof_of_work_bits)) != 0) throw CudaError("proof of work self-check failed");
        CK(cudaEventRecord(ev_[T_POW + 1], st_));
        // query rounds
        u32* h_idx = reinterpret_cast<u32*>(h_stage_ + ho_.idx);
        for (int q = 0; q < nqr; ++q) {
            u32 x = (u32)(j.ch.get() % N_);
            h_idx[q] = x;
            for (size_t i = 0; i < L; ++i) { x >>= cd_.reduction_arity_bits[i]; h_idx[(1 + i) * nqr + q] = x; }
        }
        u32* d_idx = reinterpret_cast<u32*>(query_idx_dev_.get());
        CK(cudaMemcpyAsync(d_idx, h_idx, (1 + L) * nqr * 4, cudaMemcpyHostToDevice, st_));
        u64* qo = query_out_dev_.get();
        size_t qpos = 0;
        const BatchDev* trees[4] = {&cs_, &wires_, &zs_, &quot_};
        const int plen0 = (int)(lg_N_ - cap_h);
        for (int t = 0; t < 4; ++t) {
            const int width = trees[t]->ncols + trees[t]->salt;
            j.row_off[t] = qpos;
            launch_gather_rows(trees[t]->lde.get(), N_, width, d_idx, nqr, qo + qpos, st_);
            qpos += (size_t)nqr * width;
            j.path_off[t] = qpos;
            launch_gather_paths(trees[t]->digests.get(), N_, plen0, d_idx, nqr, qo + qpos, st_);
            qpos += (size_t)nqr * plen0 * 4;
        }
        j.leaf_off.assign(L, 0); j.lpath_off.assign(L, 0); j.lplen.assign(L, 0);
        size_t mm = n_;
        unsigned bits = lg_N_;
        for (size_t i = 0; i < L; ++i) {
            const unsigned ab = (unsigned)cd_.reduction_arity_bits[i];
            const int arity = 1 << ab;
            const size_t M = mm << cd_.rate_bits;
            bits -= ab;
            j.lplen[i] = bits >= cap_h ? (int)(bits - cap_h) : 0;
            j.leaf_off[i] = qpos;
            u64* va = fri_values_[i].get();
            launch_gather_ext_leaves(va, va + M, arity, d_idx + (1 + i) * nqr, nqr, qo + qpos, st_);
            qpos += (size_t)nqr * arity * 2;
            j.lpath_off[i] = qpos;
            launch_gather_paths(fri_digests_[i].get(), M >> ab, j.lplen[i], d_idx + (1 + i) * nqr, nqr, qo + qpos, st_);
            qpos += (size_t)nqr * j.lplen[i] * 4;
            mm >>= ab;
        }
        if (qpos != q_words_) throw CudaError("internal: query staging layout mismatch");
        CK(cudaMemcpyAsync(h_stage_ + ho_.q, qo, qpos * 8, cudaMemcpyDeviceToHost, st_));
        CK(cudaEventRecord(ev_[T_QUERIES + 1], st_));
        j.stage = ST_QUERIES;
        return false;
    }
    case ST_QUERIES: {
        serialize_proof();
        // stage timings: every event was recorded before the wait that just ended
        for (int s = T_WIRES_INTT; s <= T_QUERIES; ++s) CK(cudaEventElapsedTime(&timings[s], ev_[s], ev_[s + 1]));
        CK(cudaEventElapsedTime(&timings[T_TOTAL], ev_[0], ev_[T_QUERIES + 1]));
        timings[T_HOST_TRANSCRIPT] = (float)(j.ch.seconds * 1e3);      // wall time inside the Fiat-Shamir sponge (host, serial)
        timings[T_HOST_PERMS] = (float)j.ch.permutations;
        j.stage = ST_IDLE;
        --g_proofs_in_flight;
        return true;
    }
    default: throw CudaError("internal: bad proof stage");
    }
}

// (l) ProofWithPublicInputs::to_bytes from the pinned block
void Circuit::serialize_proof() {
    Job& j = job_;
    const int ncs = cs_.ncols, nw = wires_.ncols, nzp = zs_.ncols, nq = quot_.ncols, nch = (int)cd_.num_challenges;
    const int nall = ncs + nw + nzp + nq;
    const unsigned cap_h = (unsigned)cd_.cap_height;
    const size_t cap_words = size_t(4) << cap_h, L = cd_.reduction_arity_bits.size(), fin = j.m;
    const int nqr = (int)cd_.num_query_rounds, plen0 = (int)(lg_N_ - cap_h);
    const u64* h_caps = h_stage_ + ho_.caps;
    const u64* h_open = h_stage_ + ho_.open;
    const u64* h_final = h_stage_ + ho_.fin;
    const u64* h_q = h_stage_ + ho_.q;
    const BatchDev* trees[4] = {&cs_, &wires_, &zs_, &quot_};
    ByteWriter w{j.out};
    w.words(h_caps, 3 * cap_words);                  // wires, Z / partial products, quotient caps
    const u64* o_cs = h_open;
    const u64* o_w = h_open + 2 * ncs;
    const u64* o_z = h_open + 2 * (ncs + nw);
    const u64* o_q = h_open + 2 * (ncs + nw + nzp);
    const u64* o_zn = h_open + 2 * nall;
    w.words(o_cs, 2 * (size_t)ncs);                 // constants then sigmas
    w.words(o_w, 2 * (size_t)nw);
    w.words(o_z, 2 * (size_t)nch);                  // plonk_zs
    w.words(o_zn, 2 * (size_t)nch);                 // plonk_zs_next
    w.words(o_z + 2 * nch, 2 * (size_t)(nzp - nch));  // partial products
    w.words(o_q, 2 * (size_t)nq);
    for (size_t i = 0; i < L; ++i) w.words(h_caps + (3 + i) * cap_words, cap_words);
    for (int q = 0; q < nqr; ++q) {
        for (int t = 0; t < 4; ++t) {
            const int width = trees[t]->ncols + trees[t]->salt;
            w.words(h_q + j.row_off[t] + (size_t)q * width, width);
            w.byte((uint8_t)plen0);
            w.words(h_q + j.path_off[t] + (size_t)q * plen0 * 4, (size_t)plen0 * 4);
        }
        for (size_t i = 0; i < L; ++i) {
            const size_t ar2 = size_t(2) << cd_.reduction_arity_bits[i];
            w.words(h_q + j.leaf_off[i] + (size_t)q * ar2, ar2);
            w.byte((uint8_t)j.lplen[i]);
            w.words(h_q + j.lpath_off[i] + (size_t)q * j.lplen[i] * 4, (size_t)j.lplen[i] * 4);
        }
    }
    for (size_t k = 0; k < fin; ++k) { w.w64(h_final[k]); w.w64(h_final[fin + k]); }
    w.w64(j.pow_witness);
    w.w64(j.pis.size());
    w.words(j.pis.data(), j.pis.size());
    if (w.pos != cd_.proof_size()) throw CudaError("internal: proof size mismatch");
    j.len = w.pos;
}

}  // namespace zkb
