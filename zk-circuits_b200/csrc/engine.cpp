#include "engine.hpp"
#include "../../include/zkb200.h"
#include <cstdlib>
#include <new>

namespace zkb {

int status_of_current_exception(std::string& msg) {
    try {
        throw;
    } catch (const ParseError& e) { msg = e.what(); return ZKB_E_PARSE;
    } catch (const UnsupportedError& e) { msg = e.what(); return ZKB_E_UNSUPPORTED_GATE;
    } catch (const ArgError& e) { msg = e.what(); return ZKB_E_ARG;
    } catch (const DigestError& e) { msg = e.what(); return ZKB_E_DIGEST;
    } catch (const ZetaError& e) { msg = e.what(); return ZKB_E_ZETA_IN_SUBGROUP;
    } catch (const UnsatError& e) { msg = e.what(); return ZKB_E_UNSAT;
    } catch (const BufferError& e) { msg = e.what(); return ZKB_E_BUFFER;
    } catch (const NcclError& e) { msg = e.what(); return ZKB_E_NCCL;
    } catch (const CudaError& e) { msg = e.what(); return ZKB_E_CUDA;
    } catch (const std::bad_alloc&) { msg = "out of host memory"; return ZKB_E_ARG;
    } catch (const std::exception& e) { msg = e.what(); return ZKB_E_CUDA;
    } catch (...) { msg = "unknown error"; return ZKB_E_CUDA; }
}

Engine::Engine(const uint8_t* common, size_t len, const u64* const_sigma, bool is_values, const u64* digest, int device,
               int n_contexts, int n_slots)
    : device_(device) {
    if (n_contexts < 1 || n_contexts > 64) throw ArgError("n_contexts must be in [1, 64]");
    if (n_slots < n_contexts) n_slots = n_contexts;
    if (n_slots > 256) throw ArgError("too many witness slots");
    for (int i = 0; i < n_contexts; ++i)
        ctx_.push_back(std::make_unique<Circuit>(common, len, const_sigma, is_values, digest, device));
    ctx_slot_.assign(n_contexts, -1);
    slot_.resize(n_slots);
    const CommonData& cd = ctx_[0]->common();
    const size_t bytes = (size_t)cd.num_wires * cd.degree() * sizeof(u64);
    int prev = 0;
    cuda_check(cudaGetDevice(&prev), "cudaGetDevice");
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    try {
        for (auto& s : slot_) cuda_check(cudaMallocHost(&s.wires, bytes), "cudaMallocHost(witness slot)");
    } catch (...) {
        for (auto& s : slot_) if (s.wires) cudaFreeHost(s.wires);
        cudaSetDevice(prev);
        throw;
    }
    cudaSetDevice(prev);
    // One driver thread keeps the contexts of a wormhole-sized circuit busy (a proof is ~150 launches over ~5 ms of GPU time).
    // A voting-sized proof (n <= 2^11) is the same ~130 launches over well under 1 ms of GPU time: the driver's launch calls
    // are the limit, so the contexts are split between up to four drivers (16 contexts of the n = 2^9 voting circuit: 1299 ->
    // 1678 proofs/s with four; the wormhole circuit gains 2 % and keeps one thread per GPU). ZKB_ENGINE_DRIVERS overrides (1-8).
    int n_drivers = cd.degree_bits <= 11 ? (n_contexts >= 8 ? 4 : (n_contexts >= 4 ? 2 : 1)) : 1;
    if (const char* e = std::getenv("ZKB_ENGINE_DRIVERS")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 8) n_drivers = v;
    }
    if (n_drivers > n_contexts) n_drivers = n_contexts;
    for (int t = 0; t < n_drivers; ++t) drivers_.emplace_back([this, t, n_drivers] { run(t, n_drivers); });
}

Engine::~Engine() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_driver_.notify_all();
    for (auto& d : drivers_) if (d.joinable()) d.join();
    cudaSetDevice(device_);
    for (auto& c : ctx_) if (c->busy()) c->abort_proof();
    for (auto& s : slot_) if (s.wires) cudaFreeHost(s.wires);
}

int Engine::acquire(u64** wires_buf) {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
        for (size_t i = 0; i < slot_.size(); ++i)
            if (slot_[i].state == S_FREE) {
                slot_[i].state = S_ACQUIRED;
                if (wires_buf) *wires_buf = slot_[i].wires;
                return (int)i;
            }
        cv_client_.wait(lk);
    }
}

void Engine::release(int slot) {
    std::lock_guard<std::mutex> lk(mu_);
    if (slot < 0 || slot >= (int)slot_.size() || slot_[slot].state != S_ACQUIRED) throw ArgError("slot is not held by the caller");
    slot_[slot].state = S_FREE;
    cv_client_.notify_all();
}

void Engine::submit(int slot, const u64* public_inputs, size_t n_pi, const u64* salts, u64 salt_seed, u32 flags, uint8_t* proof_out,
                    size_t proof_cap) {
    // argument errors surface here, synchronously, exactly as zkb_prove reports them
    ctx_[0]->validate_prove_args(public_inputs, n_pi, flags, proof_out, proof_cap);
    std::lock_guard<std::mutex> lk(mu_);
    if (slot < 0 || slot >= (int)slot_.size() || slot_[slot].state != S_ACQUIRED) throw ArgError("slot is not held by the caller");
    Slot& s = slot_[slot];
    s.pis.assign(public_inputs, public_inputs + n_pi);
    s.salts = salts; s.salt_seed = salt_seed; s.flags = flags; s.out = proof_out; s.cap = proof_cap;
    s.len = 0; s.status = 0; s.error.clear();
    s.state = S_QUEUED;
    queue_.push_back(slot);
    cv_driver_.notify_one();
}

int Engine::wait(int slot, size_t* proof_len, std::string* err) {
    std::unique_lock<std::mutex> lk(mu_);
    if (slot < 0 || slot >= (int)slot_.size()) throw ArgError("bad slot");
    Slot& s = slot_[slot];
    if (s.state != S_QUEUED && s.state != S_RUNNING && s.state != S_DONE) throw ArgError("nothing was submitted on this slot");
    cv_client_.wait(lk, [&] { return s.state == S_DONE; });
    if (proof_len) *proof_len = s.len;
    if (err) *err = s.error;
    const int rc = s.status;
    s.state = S_FREE;
    cv_client_.notify_all();
    return rc;
}

void Engine::finish(int c, int status, const std::string& err) {
    std::lock_guard<std::mutex> lk(mu_);
    Slot& s = slot_[ctx_slot_[c]];
    s.status = status;
    s.error = err;
    s.len = status == 0 ? ctx_[c]->proof_len() : 0;
    s.state = S_DONE;
    ctx_slot_[c] = -1;
    cv_client_.notify_all();
}

// The driver: start queued proofs on idle contexts, advance whichever context's stream has drained. It spins while proofs
// are in flight (one thread per GPU; the stage boundaries are 0.1-2 ms apart) and sleeps on the condition variable otherwise.
void Engine::run(int driver, int n_drivers) {
    cudaSetDevice(device_);
    const int nc = (int)ctx_.size();
    unsigned idle_spins = 0;
    for (;;) {
        bool progressed = false;
        int running = 0;
        for (int c = driver; c < nc; c += n_drivers) {
            if (ctx_slot_[c] < 0) {
                int slot = -1;
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    if (!queue_.empty()) { slot = queue_.front(); queue_.pop_front(); slot_[slot].state = S_RUNNING; ctx_slot_[c] = slot; }
                }
                if (slot < 0) continue;
                Slot& s = slot_[slot];
                try {
                    if (!(s.flags & PF_WITNESS_RESIDENT)) ctx_[c]->upload_witness(s.wires, /*wait=*/false);
                    ctx_[c]->begin_proof(s.pis.data(), s.pis.size(), s.salts, s.salt_seed, s.flags & ~PF_WITNESS_RESIDENT, s.out, s.cap);
                } catch (...) {
                    std::string msg;
                    const int rc = status_of_current_exception(msg);
                    ctx_[c]->abort_proof();
                    finish(c, rc, msg);
                }
                progressed = true;
                continue;
            }
            ++running;
            try {
                if (!ctx_[c]->ready()) continue;
                progressed = true;
                if (ctx_[c]->advance()) finish(c, 0, std::string());
            } catch (...) {
                std::string msg;
                const int rc = status_of_current_exception(msg);
                ctx_[c]->abort_proof();
                finish(c, rc, msg);
            }
        }
        if (progressed) { idle_spins = 0; continue; }
        if (running) {
            if (++idle_spins > 64) std::this_thread::yield();
            continue;
        }
        std::unique_lock<std::mutex> lk(mu_);
        if (stop_) return;
        if (queue_.empty()) cv_driver_.wait(lk, [&] { return stop_ || !queue_.empty(); });
        if (stop_ && queue_.empty()) return;
    }
}

}  // namespace zkb
