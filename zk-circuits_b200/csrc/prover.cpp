#include "prover.hpp"
#include "host_transcript.hpp"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <thread>

namespace zkb {

static std::atomic<int> g_live_contexts{0};     // prover contexts alive in this process (see Circuit::sync)

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
    // ZKB_TRACE=1: wall-clock checkpoints of the context build on stderr (which part of zkb_circuit_create costs what)
    const bool trace = std::getenv("ZKB_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (trace) std::fprintf(stderr, "[zkb create] %-28s %8.3f ms\n", what,
                                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) throw ArgError("bad device index");
    DeviceGuard g(device);
    device_tables_init(device);
    CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    for (auto& e : ev_) CK(cudaEventCreate(&e));
    CK(cudaEventCreateWithFlags(&sync_ev_, cudaEventBlockingSync | cudaEventDisableTiming));
    ++g_live_contexts;
    lg_n_ = (unsigned)cd_.degree_bits;
    lg_N_ = lg_n_ + (unsigned)cd_.rate_bits;
    n_ = size_t(1) << lg_n_;
    N_ = size_t(1) << lg_N_;
    const int ncs = (int)(cd_.num_constants + cd_.num_routed_wires), nw = (int)cd_.num_wires;
    const int nzp = (int)cd_.num_zs_pp(), nq = (int)cd_.num_quotient_polys(), nch = (int)cd_.num_challenges;
    const int salt = (int)cd_.salt_size();
    const unsigned cap_h = (unsigned)cd_.cap_height;
    mark("stream, events, tables");
    check_canonical(const_sigma, (size_t)ncs * n_, "const_sigma");
    mark("host canonical check");

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
    want(sigma_vals_, (size_t)cd_.num_routed_wires * n_);
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
    want(query_idx_dev_, ((1 + L) * nqr + 1) / 2 + 1);
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
    h_stage_words_ = qwords + 4096 + 2 * (size_t)(nall + nch) + 2 * cd_.final_poly_len() + 2 * nall;
    h_qp_off_ = h_stage_words_;     // separate region for the quotient parameters + alpha powers
    h_stage_words_ += (sizeof(QuotientParams) + 7) / 8 + (size_t)nch * nterms + 8;
    mark("device buffers");
    CK(cudaMallocHost(&h_stage_, h_stage_words_ * sizeof(u64)));
    mark("pinned staging");

    CK(cudaMemcpyAsync(k_is_dev_.get(), cd_.k_is.data(), cd_.k_is.size() * 8, cudaMemcpyHostToDevice, st_));
    // ---- constants/sigmas commitment (the part of CircuitBuilder::build the prover needs) ----
    const size_t nconst = cd_.num_constants;
    if (is_values) {
        CK(cudaMemcpyAsync(cs_.coeff_ptr, const_sigma, (size_t)ncs * n_ * 8, cudaMemcpyHostToDevice, st_));
        CK(cudaMemcpyAsync(sigma_vals_.get(), cs_.coeff_ptr + nconst * n_, cd_.num_routed_wires * n_ * 8, cudaMemcpyDeviceToDevice, st_));
        launch_intt_natural(cs_.coeff_ptr, n_, cs_.coeff_ptr, n_, ncs, lg_n_, nullptr, st_);
    } else {
        CK(cudaMemcpyAsync(cs_.coeff_ptr, const_sigma, (size_t)ncs * n_ * 8, cudaMemcpyHostToDevice, st_));
        // sigma values over H: evaluate on <w_n> (rate 0, shift 1) then undo the leaf ordering
        launch_lde(cs_.coeff_ptr + nconst * n_, n_, sigma_vals_.get(), n_, (int)cd_.num_routed_wires, lg_n_, 0, 1, st_);
        launch_bitrev_permute(sigma_vals_.get(), n_, (int)cd_.num_routed_wires, lg_n_, st_);
    }
    cs_cap_.resize((size_t(4)) << cap_h);
    commit_batch(cs_, 0, nullptr, 0, h_stage_);
    mark("uploads and launches queued");
    sync();
    mark("constants/sigmas commitment");
    std::memcpy(cs_cap_.data(), h_stage_, cs_cap_.size() * 8);
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

Circuit::~Circuit() {
    cudaSetDevice(device_);
    if (st_) cudaStreamSynchronize(st_);
    for (auto& kv : level_graphs_) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    for (auto& e : ev_) if (e) cudaEventDestroy(e);
    for (auto& a : qfork_.aux) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
    if (qfork_.fork) cudaEventDestroy(qfork_.fork);
    for (auto& e : qfork_.join) if (e) cudaEventDestroy(e);
    if (sync_ev_) { cudaEventDestroy(sync_ev_); --g_live_contexts; }
    if (h_stage_) cudaFreeHost(h_stage_);
    if (st_) cudaStreamDestroy(st_);
}

// Host waits on the stream ~10 times per proof. Spinning (the runtime's default) gives the lowest latency. With more
// proofs in flight in a process than host cores available to it (8 ranks x 8 streams on a 32-core box) the wait polls
// every ~20 us and sleeps in between instead, so that the threads doing Fiat-Shamir work are not starved by spinners:
// measured at 8 GPUs / 32 cores, 8 streams per GPU sleep-polling 1622 proofs/s (98 % of 8 x one GPU), 4 streams per GPU
// spinning 1419, 8 streams yielding 1410 (profiles/r01_bench_n8_sync_modes.json). A blocking (interrupt) wait is much
// worse (708). ZKB_SYNC=spin|yield|block|sleep overrides the choice.
static std::atomic<int> g_proofs_in_flight{0};
struct InFlight {
    InFlight() { ++g_proofs_in_flight; }
    ~InFlight() { --g_proofs_in_flight; }
};
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

// LDE (+ salts) + Merkle for a batch whose coefficients are in place; cap is copied to cap_host (pinned) asynchronously
void Circuit::commit_batch(BatchDev& b, unsigned batch_id, const u64* salts_host, u64 salt_seed, u64* cap_host) {
    const unsigned cap_h = (unsigned)cd_.cap_height;
    launch_lde(b.coeff_ptr, b.coeff_stride, b.lde.get(), N_, b.ncols, lg_n_, (unsigned)cd_.rate_bits, GL_GEN, st_);
    if (b.salt) {
        u64* sp = b.lde.get() + (size_t)b.ncols * N_;
        if (salts_host) CK(cudaMemcpyAsync(sp, salts_host, (size_t)b.salt * N_ * 8, cudaMemcpyHostToDevice, st_));
        else launch_salt_fill(sp, N_, N_, salt_seed, batch_id, st_);
    }
    launch_merkle_leaves(b.lde.get(), N_, b.ncols + b.salt, N_, b.digests.get(), st_);
    b.cap_offset = run_merkle_levels(b.digests.get(), N_, cap_h);
    CK(cudaMemcpyAsync(cap_host, b.digests.get() + b.cap_offset * 4, (size_t(32)) << cap_h, cudaMemcpyDeviceToHost, st_));
}

// H2D of the wire matrix + canonical check on the device. With wait = false nothing synchronises: the copy and the check
// are queued ahead of the proof's kernels and the flag is read at the proof's first sync point (zkb_prove path).
void Circuit::upload_witness(const u64* wires_host, bool wait) {
    if (!wires_host) throw ArgError("wires is null");
    DeviceGuard g(device_);
    unsigned* flag = reinterpret_cast<unsigned*>(flag_dev_.get());
    CK(cudaMemsetAsync(flag, 0, sizeof(u64), st_));
    CK(cudaMemcpyAsync(wires_vals_.get(), wires_host, cd_.num_wires * n_ * 8, cudaMemcpyHostToDevice, st_));
    launch_canonical_check(wires_vals_.get(), cd_.num_wires * n_, flag, st_);
    check_pending_ = true;
    if (wait) {
        sync();
        finish_witness_check();
    }
}
void Circuit::finish_witness_check() {     // stream must be idle
    if (!check_pending_) return;
    check_pending_ = false;
    u64 f = 0;
    CK(cudaMemcpy(&f, flag_dev_.get(), sizeof(u64), cudaMemcpyDeviceToHost));
    if (f) throw ArgError("wires: non-canonical field element");
}

void Circuit::run_partial_products(const u64* betas, const u64* gammas) {
    u64 bg[8];
    const int nch = (int)cd_.num_challenges;
    for (int c = 0; c < nch; ++c) { bg[c] = betas[c]; bg[nch + c] = gammas[c]; }
    launch_partial_products(wires_vals_.get(), n_, sigma_vals_.get(), n_, k_is_dev_.get(), (int)cd_.num_routed_wires,
                            (int)cd_.quotient_degree_factor, nch, bg, lg_n_, zs_vals_.get(), n_, pp_scratch_.get(), st_);
}

void Circuit::run_quotient(const u64* pi_hash, const u64* betas, const u64* gammas, const u64* alphas) {
    const int nch = (int)cd_.num_challenges;
    const int nterms = nch * (2 + (int)cd_.num_partial_products) + (int)cd_.num_gate_constraints;
    QuotientParams* qp = reinterpret_cast<QuotientParams*>(h_stage_ + h_qp_off_);
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
    u64* ap = h_stage_ + h_qp_off_ + (sizeof(QuotientParams) + 7) / 8;
    for (int c = 0; c < nch; ++c) {
        u64 p = 1;
        for (int k = 0; k < nterms; ++k) { ap[(size_t)c * nterms + k] = p; p = gl_mul(p, alphas[c]); }
    }
    CK(cudaMemcpyAsync(qparams_dev_.get(), qp, sizeof(QuotientParams), cudaMemcpyHostToDevice, st_));
    CK(cudaMemcpyAsync(apow_dev_.get(), ap, (size_t)nch * nterms * 8, cudaMemcpyHostToDevice, st_));
    launch_quotient(reinterpret_cast<const QuotientParams*>(qparams_dev_.get()), *qp, apow_dev_.get(), nterms, cs_.lde.get(), N_,
                    wires_.lde.get(), N_, zs_.lde.get(), N_, q_.get(), N_, st_, q_sliced_ ? &qfork_ : nullptr);
    // values on the coset (leaf order) -> coefficients; chunks of n are the committed quotient polynomials
    launch_coset_intt_bitrev(q_.get(), N_, nch, lg_N_, GL_GEN, st_);
}

void Circuit::partial_products(const u64* wires_host, const u64* betas, const u64* gammas, u64* out_host) {
    if (!wires_host || !betas || !gammas || !out_host) throw ArgError("null argument");
    DeviceGuard g(device_);
    CK(cudaMemcpyAsync(wires_vals_.get(), wires_host, cd_.num_wires * n_ * 8, cudaMemcpyHostToDevice, st_));
    run_partial_products(betas, gammas);
    CK(cudaMemcpyAsync(out_host, zs_vals_.get(), cd_.num_zs_pp() * n_ * 8, cudaMemcpyDeviceToHost, st_));
    sync();
}

void Circuit::quotient(const u64* wires_host, const u64* zs_pp_host, const u64* pis, size_t n_pi, const u64* betas,
                       const u64* gammas, const u64* alphas, u64* out_host) {
    if (!wires_host || !zs_pp_host || !betas || !gammas || !alphas || !out_host) throw ArgError("null argument");
    DeviceGuard g(device_);
    const int nw = (int)cd_.num_wires, nzp = (int)cd_.num_zs_pp();
    CK(cudaMemcpyAsync(wires_vals_.get(), wires_host, (size_t)nw * n_ * 8, cudaMemcpyHostToDevice, st_));
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
}  // namespace

size_t Circuit::prove_resident(const u64* public_inputs, size_t n_pi, const u64* salts, u64 salt_seed, u32 pow_rule,
                               uint8_t* out, size_t cap) {
    if (n_pi != cd_.num_public_inputs) throw ArgError("wrong number of public inputs");
    if (n_pi && !public_inputs) throw ArgError("public_inputs is null");
    if (pow_rule != 0) throw ArgError("unknown pow_rule");
    const size_t psize = cd_.proof_size();
    if (!out || cap < psize) throw BufferError(psize);
    check_canonical(public_inputs, n_pi, "public_inputs");
    DeviceGuard g(device_);
    InFlight in_flight;

    const int ncs = cs_.ncols, nw = wires_.ncols, nzp = zs_.ncols, nq = quot_.ncols, nch = (int)cd_.num_challenges;
    const int nall = ncs + nw + nzp + nq;
    const unsigned cap_h = (unsigned)cd_.cap_height, rate_bits = (unsigned)cd_.rate_bits;
    const size_t cap_words = size_t(4) << cap_h;
    const size_t L = cd_.reduction_arity_bits.size();
    const int nqr = (int)cd_.num_query_rounds;
    const u64* salt_ptr[3] = {nullptr, nullptr, nullptr};
    if (salts && cd_.salt_size())
        for (int b = 0; b < 3; ++b) salt_ptr[b] = salts + (size_t)b * 4 * N_;

    // pinned staging layout
    u64* h_caps = h_stage_;                         // up to (3 + L) caps
    u64* h_open = h_caps + (3 + L) * cap_words;     // 2 * (nall + nch)
    u64* h_misc = h_open + 2 * (size_t)(nall + nch);

    u64 pi_hash[4];
    h_hash_no_pad(public_inputs, n_pi, pi_hash);
    Challenger ch;
    ch.observe_many(circuit_digest_, 4);
    ch.observe_many(pi_hash, 4);

    CK(cudaEventRecord(ev_[0], st_));
    // (d) wires commitment
    launch_intt_natural(wires_vals_.get(), n_, wires_.coeff_ptr, n_, nw, lg_n_, nullptr, st_);
    CK(cudaEventRecord(ev_[T_WIRES_INTT + 1], st_));
    launch_lde(wires_.coeff_ptr, n_, wires_.lde.get(), N_, nw, lg_n_, rate_bits, GL_GEN, st_);
    CK(cudaEventRecord(ev_[T_WIRES_LDE + 1], st_));      // exactly the coset-LDE launch: bench.py's roofline kernel
    if (wires_.salt) {
        u64* sp = wires_.lde.get() + (size_t)nw * N_;
        if (salt_ptr[0]) CK(cudaMemcpyAsync(sp, salt_ptr[0], 4 * N_ * 8, cudaMemcpyHostToDevice, st_));
        else launch_salt_fill(sp, N_, N_, salt_seed, 0, st_);
    }
    launch_merkle_leaves(wires_.lde.get(), N_, nw + wires_.salt, N_, wires_.digests.get(), st_);
    wires_.cap_offset = run_merkle_levels(wires_.digests.get(), N_, cap_h);
    u64* wires_cap = h_caps;
    CK(cudaMemcpyAsync(wires_cap, wires_.digests.get() + wires_.cap_offset * 4, cap_words * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaEventRecord(ev_[T_WIRES_MERKLE + 1], st_));
    sync();
    finish_witness_check();
    ch.observe_many(wires_cap, cap_words);
    u64 betas[2], gammas[2], alphas[2];
    for (int c = 0; c < nch; ++c) betas[c] = ch.get();
    for (int c = 0; c < nch; ++c) gammas[c] = ch.get();

    // (f) partial products and Z, (g) commit
    run_partial_products(betas, gammas);
    CK(cudaEventRecord(ev_[T_PP + 1], st_));
    launch_intt_natural(zs_vals_.get(), n_, zs_vals_.get(), n_, nzp, lg_n_, nullptr, st_);
    u64* zs_cap = h_caps + cap_words;
    commit_batch(zs_, 1, salt_ptr[1], salt_seed, zs_cap);
    CK(cudaEventRecord(ev_[T_ZS_COMMIT + 1], st_));
    sync();
    ch.observe_many(zs_cap, cap_words);
    for (int c = 0; c < nch; ++c) alphas[c] = ch.get();

    // (h) quotient, (i) commit
    run_quotient(pi_hash, betas, gammas, alphas);
    CK(cudaEventRecord(ev_[T_QUOTIENT + 1], st_));
    u64* quot_cap = h_caps + 2 * cap_words;
    commit_batch(quot_, 2, salt_ptr[2], salt_seed, quot_cap);
    CK(cudaEventRecord(ev_[T_QUOTIENT_COMMIT + 1], st_));
    sync();
    ch.observe_many(quot_cap, cap_words);
    ext2 zeta = ch.get_ext();
    if (e_eq(e_pow2k(zeta, lg_n_), e_from(1))) throw ZetaError("Opening point is in the subgroup.");
    ext2 zeta_next = e_mul_base(zeta, gl_root_of_unity(lg_n_));

    // (j) openings
    u64* od = openings_dev_.get();
    {
        OpeningsArgs oa;
        oa.seg[0] = {cs_.coeff_ptr, n_, ncs, 0};
        oa.seg[1] = {wires_.coeff_ptr, n_, nw, 0};
        oa.seg[2] = {zs_.coeff_ptr, n_, nzp, 0};
        oa.seg[3] = {quot_.coeff_ptr, n_, nq, 0};
        oa.seg[4] = {zs_.coeff_ptr, n_, nch, 1};          // Z polynomials again, at g * zeta
        oa.nseg = 5; oa.lg_n = lg_n_; oa.pw = zpow_.get(); oa.out = od;
        launch_openings(oa, zeta, zeta_next, st_);
    }
    CK(cudaMemcpyAsync(h_open, od, 2 * (size_t)(nall + nch) * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaEventRecord(ev_[T_OPENINGS + 1], st_));
    sync();
    // transcript order: constants, sigmas, wires, zs, partial products, quotient, then zs_next — which is the
    // device order (cs, wires, zs_pp, quot, zs_next)
    ch.observe_many(h_open, 2 * (size_t)(nall + nch));

    // (k) prove_openings: batch combination -> final polynomial
    ext2 alpha = ch.get_ext();
    u64* h_apow = h_misc;                       // SoA: a[nall], b[nall]
    ext2 reduced0 = e_make(0, 0), reduced1 = e_make(0, 0);
    {
        ext2 p = e_from(1);
        for (int j = 0; j < nall; ++j) {
            h_apow[j] = p.a;
            h_apow[nall + j] = p.b;
            reduced0 = e_add(reduced0, e_mul(p, e_make(h_open[2 * j], h_open[2 * j + 1])));
            if (j < nch) reduced1 = e_add(reduced1, e_mul(p, e_make(h_open[2 * (nall + j)], h_open[2 * (nall + j) + 1])));
            p = e_mul(p, alpha);
        }
    }
    CK(cudaMemcpyAsync(fri_apow_.get(), h_apow, 2 * (size_t)nall * 8, cudaMemcpyHostToDevice, st_));
    FriCombineParams fp;
    fp.lde[0] = cs_.lde.get(); fp.lde[1] = wires_.lde.get(); fp.lde[2] = zs_.lde.get(); fp.lde[3] = quot_.lde.get();
    for (int t = 0; t < 4; ++t) fp.stride[t] = N_;
    fp.ncols[0] = ncs; fp.ncols[1] = nw; fp.ncols[2] = nzp; fp.ncols[3] = nq;
    fp.num_zs = nch;
    fp.alpha = alpha; fp.zeta = zeta; fp.zeta_next = zeta_next; fp.reduced0 = reduced0; fp.reduced1 = reduced1;
    fp.lg_n = lg_n_;
    u64* fc = fri_coeffs_[0].get();
    launch_fri_combine(fp, fri_apow_.get(), fri_apow_.get() + nall, fc, fc + n_, st_);
    launch_coset_intt_bitrev(fc, n_, 2, lg_n_, GL_GEN, st_);
    CK(cudaEventRecord(ev_[T_FRI_COMBINE + 1], st_));

    // FRI commit phase
    u64 shift = GL_GEN;
    size_t m = n_;
    unsigned lg_m = lg_n_;
    std::vector<ext2> fri_betas;
    for (size_t i = 0; i < L; ++i) {
        const unsigned ab = (unsigned)cd_.reduction_arity_bits[i];
        const int arity = 1 << ab;
        const size_t M = m << rate_bits;
        u64* ca = fri_coeffs_[i].get();
        u64* va = fri_values_[i].get();
        launch_lde(ca, m, va, M, 2, lg_m, rate_bits, shift, st_);
        const size_t leaves = M >> ab;
        launch_merkle_leaves_ext(va, va + M, arity, leaves, fri_digests_[i].get(), st_);
        fri_cap_off_[i] = run_merkle_levels(fri_digests_[i].get(), leaves, cap_h);
        u64* lcap = h_caps + (3 + i) * cap_words;
        CK(cudaMemcpyAsync(lcap, fri_digests_[i].get() + fri_cap_off_[i] * 4, cap_words * 8, cudaMemcpyDeviceToHost, st_));
        sync();
        ch.observe_many(lcap, cap_words);
        ext2 beta = ch.get_ext();
        fri_betas.push_back(beta);
        u64* na = fri_coeffs_[i + 1].get();
        launch_fri_fold(ca, ca + m, na, na + (m >> ab), m >> ab, arity, beta, st_);
        m >>= ab;
        lg_m -= ab;
        shift = gl_pow(shift, (u64)arity);
    }
    const size_t fin = m;   // == final_poly_len
    u64* h_final = h_misc + 2 * (size_t)nall;
    CK(cudaMemcpyAsync(h_final, fri_coeffs_[L].get(), 2 * fin * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaEventRecord(ev_[T_FRI_COMMIT + 1], st_));
    sync();
    for (size_t k = 0; k < fin; ++k) ch.observe_ext(e_make(h_final[k], h_final[fin + k]));

    // proof of work (MIN rule): smallest witness whose response has >= pow_bits leading zeros
    u64 pow_witness = 0;
    {
        u64* h_pow = h_final + 2 * fin;
        for (int i = 0; i < 12; ++i) h_pow[i] = ch.sponge[i];
        const int pos = ch.in_len;
        for (int i = 0; i < pos; ++i) h_pow[i] = ch.in_buf[i];
        h_pow[12] = ~u64(0);
        CK(cudaMemcpyAsync(pow_dev_.get(), h_pow, 13 * 8, cudaMemcpyHostToDevice, st_));
        const u64 batch = u64(1) << 32;   // one persistent launch scans this range in increasing order and stops at the minimum
        u64 base = 0;
        for (;;) {
            launch_pow_search(pow_dev_.get(), pos, base, batch, cd_.proof_of_work_bits,
                              reinterpret_cast<unsigned long long*>(pow_dev_.get() + 12), st_);
            CK(cudaMemcpyAsync(h_pow + 13, pow_dev_.get() + 12, 8, cudaMemcpyDeviceToHost, st_));
            sync();
            if (h_pow[13] != ~u64(0)) { pow_witness = h_pow[13]; break; }
            base += batch;
            if (base >= GL_P - batch) throw CudaError("proof of work search exhausted");
        }
        ch.observe(pow_witness);
        u64 resp = ch.get();
        if (cd_.proof_of_work_bits && (resp >> (64 - cd_.proof_of_work_bits)) != 0) throw CudaError("proof of work self-check failed");
    }
    CK(cudaEventRecord(ev_[T_POW + 1], st_));

    // query rounds
    u32* h_idx = reinterpret_cast<u32*>(h_final + 2 * fin + 16);
    for (int q = 0; q < nqr; ++q) {
        u32 x = (u32)(ch.get() % N_);
        h_idx[q] = x;
        for (size_t i = 0; i < L; ++i) { x >>= cd_.reduction_arity_bits[i]; h_idx[(1 + i) * nqr + q] = x; }
    }
    u32* d_idx = reinterpret_cast<u32*>(query_idx_dev_.get());
    CK(cudaMemcpyAsync(d_idx, h_idx, (1 + L) * nqr * 4, cudaMemcpyHostToDevice, st_));
    u64* qo = query_out_dev_.get();
    size_t qpos = 0;
    const BatchDev* trees[4] = {&cs_, &wires_, &zs_, &quot_};
    const int plen0 = (int)(lg_N_ - cap_h);
    size_t row_off[4], path_off[4];
    for (int t = 0; t < 4; ++t) {
        const int width = trees[t]->ncols + trees[t]->salt;
        row_off[t] = qpos;
        launch_gather_rows(trees[t]->lde.get(), N_, width, d_idx, nqr, qo + qpos, st_);
        qpos += (size_t)nqr * width;
        path_off[t] = qpos;
        launch_gather_paths(trees[t]->digests.get(), N_, plen0, d_idx, nqr, qo + qpos, st_);
        qpos += (size_t)nqr * plen0 * 4;
    }
    std::vector<size_t> leaf_off(L), lpath_off(L);
    std::vector<int> lplen(L);
    {
        size_t mm = n_;
        unsigned bits = lg_N_;
        for (size_t i = 0; i < L; ++i) {
            const unsigned ab = (unsigned)cd_.reduction_arity_bits[i];
            const int arity = 1 << ab;
            const size_t M = mm << rate_bits;
            bits -= ab;
            lplen[i] = bits >= cap_h ? (int)(bits - cap_h) : 0;
            leaf_off[i] = qpos;
            u64* va = fri_values_[i].get();
            launch_gather_ext_leaves(va, va + M, arity, d_idx + (1 + i) * nqr, nqr, qo + qpos, st_);
            qpos += (size_t)nqr * arity * 2;
            lpath_off[i] = qpos;
            launch_gather_paths(fri_digests_[i].get(), M >> ab, lplen[i], d_idx + (1 + i) * nqr, nqr, qo + qpos, st_);
            qpos += (size_t)nqr * lplen[i] * 4;
            mm >>= ab;
        }
    }
    u64* h_q = h_final + 2 * fin + 16 + ((1 + L) * nqr + 1) / 2 + 1;
    if ((size_t)(h_q - h_stage_) + qpos > h_stage_words_) throw CudaError("internal: staging buffer too small");
    CK(cudaMemcpyAsync(h_q, qo, qpos * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaEventRecord(ev_[T_QUERIES + 1], st_));
    sync();

    // (l) serialise: ProofWithPublicInputs::to_bytes
    ByteWriter w{out};
    w.words(wires_cap, cap_words);
    w.words(zs_cap, cap_words);
    w.words(quot_cap, cap_words);
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
            w.words(h_q + row_off[t] + (size_t)q * width, width);
            w.byte((uint8_t)plen0);
            w.words(h_q + path_off[t] + (size_t)q * plen0 * 4, (size_t)plen0 * 4);
        }
        for (size_t i = 0; i < L; ++i) {
            const size_t ar2 = size_t(2) << cd_.reduction_arity_bits[i];
            w.words(h_q + leaf_off[i] + (size_t)q * ar2, ar2);
            w.byte((uint8_t)lplen[i]);
            w.words(h_q + lpath_off[i] + (size_t)q * lplen[i] * 4, (size_t)lplen[i] * 4);
        }
    }
    for (size_t k = 0; k < fin; ++k) { w.w64(h_final[k]); w.w64(h_final[fin + k]); }
    w.w64(pow_witness);
    w.w64(n_pi);
    w.words(public_inputs, n_pi);
    if (w.pos != psize) throw CudaError("internal: proof size mismatch");

    // stage timings
    CK(cudaEventRecord(ev_[T_TOTAL + 1], st_));
    sync();
    for (int s = T_WIRES_INTT; s <= T_QUERIES; ++s) CK(cudaEventElapsedTime(&timings[s], ev_[s], ev_[s + 1]));
    CK(cudaEventElapsedTime(&timings[T_TOTAL], ev_[0], ev_[T_QUERIES + 1]));
    timings[T_HOST_TRANSCRIPT] = (float)(ch.seconds * 1e3);      // wall time inside the Fiat-Shamir sponge (host, serial)
    timings[T_HOST_PERMS] = (float)ch.permutations;
    return psize;
}

}  // namespace zkb
