// Proof engine: several prover contexts of ONE circuit on ONE GPU, driven by a single host thread (up to four for small
// circuits, whose proofs are launch-bound: see Engine::Engine), behind an asynchronous
// submit / wait interface with pinned witness buffers the caller fills directly.
//
// Why (SURVEY.md §8f rank 3, and the host side of §8e(1)): a proof is ~10 short GPU stages separated by serial Fiat-Shamir
// steps on the host. With one blocking host thread per proof in flight (the reference's rayon callers,
// /root/reference/wormhole/aggregator/src/circuits/tree.rs:93-103) an 8-GPU box runs 64 waiting threads on 32 cores. Here a
// proof is a resumable job (Circuit::begin_proof / ready / advance): the driver thread polls the contexts' streams, runs the
// transcript step of whichever is ready and queues its next stage, so one thread keeps all contexts of a GPU busy and the
// caller's threads are free to generate the next witnesses (/root/reference/wormhole/prover/src/lib.rs:209-225 `commit`
// + the generator graph inside `prove`) straight into the engine's pinned slots while earlier proofs run.
#pragma once
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "prover.hpp"

namespace zkb {

class Engine {
public:
    // n_contexts prover contexts (device memory: one circuit context each) and n_slots >= n_contexts pinned witness slots
    Engine(const uint8_t* common, size_t len, const u64* const_sigma, bool is_values, const u64* digest, int device,
           int n_contexts, int n_slots);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    const CommonData& common() const { return ctx_[0]->common(); }
    int contexts() const { return (int)ctx_.size(); }
    int slots() const { return (int)slot_.size(); }
    // blocks until a witness slot is free; *wires_buf = pinned [num_wires][n] buffer to fill. Returns the slot id.
    int acquire(u64** wires_buf);
    // queue the proof of the witness in `slot`. public_inputs are copied; salts (if any) and proof_out must stay valid
    // until wait() returns. With PF_WITNESS_RESIDENT the slot's buffer is ignored and the context re-proves the witness it
    // already holds (benchmark: device-resident arm; every context must have been loaded with the same witness).
    void submit(int slot, const u64* public_inputs, size_t n_pi, const u64* salts, u64 salt_seed, u32 flags, uint8_t* proof_out,
                size_t proof_cap);
    // blocks until the slot's proof is done, releases the slot. Returns a zkb_status-compatible code; *err gets the message.
    int wait(int slot, size_t* proof_len, std::string* err);
    // give a filled / acquired slot back without proving
    void release(int slot);

private:
    enum SlotState { S_FREE = 0, S_ACQUIRED, S_QUEUED, S_RUNNING, S_DONE };
    struct Slot {
        u64* wires = nullptr;           // pinned
        SlotState state = S_FREE;
        std::vector<u64> pis;
        const u64* salts = nullptr;
        u64 salt_seed = 0;
        u32 flags = 0;
        uint8_t* out = nullptr;
        size_t cap = 0, len = 0;
        int status = 0;
        std::string error;
    };
    void run(int driver, int n_drivers);
    void finish(int ctx, int status, const std::string& err);

    int device_;
    std::vector<std::unique_ptr<Circuit>> ctx_;
    std::vector<int> ctx_slot_;         // slot a context is proving, or -1
    std::vector<Slot> slot_;
    std::deque<int> queue_;
    std::mutex mu_;
    std::condition_variable cv_driver_, cv_client_;
    bool stop_ = false;
    std::vector<std::thread> drivers_;   // driver t owns contexts t, t + n_drivers, ...
};

// maps the library's exceptions to zkb_status codes (shared with capi.cpp)
int status_of_current_exception(std::string& message);

}  // namespace zkb
