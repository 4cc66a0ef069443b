// C ABI of libzkb200.so (include/zkb200.h). No exception crosses this boundary; errors map to zkb_status.
#include "../../include/zkb200.h"
#include "prover.hpp"
#include "engine.hpp"
#include "sharded.hpp"
#include <cstring>
#include <new>

using namespace zkb;

struct zkb_circuit {
    std::unique_ptr<Circuit> impl;
};
struct zkb_engine {
    std::unique_ptr<Engine> impl;
};
struct zkb_comm {
    std::unique_ptr<Comm> impl;
};

namespace {
thread_local std::string g_last_error;

template <class F>
int guarded(F&& f) {
    try {
        return f();
    } catch (...) {
        return status_of_current_exception(g_last_error);
    }
}

void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        throw CudaError(std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) throw ArgError("bad device index");
    cudaDeviceProp prop;
    cuda_check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    if (prop.major != 10) throw CudaError("device is not sm_100 (B200); the kernels are built for sm_100a only");
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    device_tables_init(device);
}
void check_canon(const uint64_t* v, size_t n, const char* what) {
    for (size_t i = 0; i < n; ++i)
        if (v[i] >= GL_P) throw ArgError(std::string(what) + ": non-canonical field element at index " + std::to_string(i));
}
bool is_pow2(size_t x) { return x && !(x & (x - 1)); }
unsigned lg2(size_t x) { unsigned k = 0; while ((size_t(1) << k) < x) ++k; return k; }
}  // namespace

extern "C" {

const char* zkb_version(void) { return "zkb200 0.2.0 (sm_100a)"; }
const char* zkb_last_error(void) { return g_last_error.c_str(); }
unsigned long long zkb_kernel_launch_count(void) { return kernel_launch_count(); }
int zkb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int zkb_circuit_create(const uint8_t* common_bin, size_t common_len, const uint64_t* const_sigma, int is_values,
                       const uint64_t circuit_digest[4], int device, zkb_circuit** out) {
    return guarded([&] {
        if (!out) throw ArgError("out is null");
        *out = nullptr;
        if (!common_bin || !const_sigma) throw ArgError("null argument");
        // parse first so that malformed inputs are reported as such even without a GPU
        (void)parse_common_data(common_bin, common_len);
        require_device(device);
        auto c = std::make_unique<zkb_circuit>();
        c->impl = std::make_unique<Circuit>(common_bin, common_len, const_sigma, is_values != 0, circuit_digest, device);
        *out = c.release();
        return (int)ZKB_OK;
    });
}
int zkb_circuit_destroy(zkb_circuit* c) {
    return guarded([&] { delete c; return (int)ZKB_OK; });
}
int zkb_circuit_verifier_only(const zkb_circuit* c, uint64_t* cap_out, size_t cap_words, uint64_t digest_out[4]) {
    return guarded([&] {
        if (!c) throw ArgError("circuit is null");
        if (cap_out && cap_words < (size_t(4) << c->impl->common().cap_height)) throw BufferError(size_t(4) << c->impl->common().cap_height);
        c->impl->verifier_only(cap_out, digest_out);
        return (int)ZKB_OK;
    });
}
size_t zkb_proof_size(const zkb_circuit* c) { return c ? c->impl->common().proof_size() : 0; }

int zkb_witness_upload(zkb_circuit* c, const uint64_t* wires) {
    return guarded([&] {
        if (!c) throw ArgError("circuit is null");
        c->impl->upload_witness(wires);
        return (int)ZKB_OK;
    });
}
int zkb_prove_resident(zkb_circuit* c, const uint64_t* public_inputs, size_t n_pi, const uint64_t* salts, uint64_t salt_seed,
                       uint32_t flags, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    int rc = guarded([&] {
        if (!c) throw ArgError("circuit is null");
        if (flags & PF_WITNESS_RESIDENT) throw ArgError("ZKB_WITNESS_RESIDENT is an engine flag");
        size_t n = c->impl->prove_resident(public_inputs, n_pi, salts, salt_seed, flags, proof_out, proof_cap);
        if (proof_len) *proof_len = n;
        return (int)ZKB_OK;
    });
    if (rc == ZKB_E_BUFFER && proof_len && c) *proof_len = c->impl->common().proof_size();
    return rc;
}
int zkb_prove(zkb_circuit* c, const uint64_t* wires, const uint64_t* public_inputs, size_t n_pi, const uint64_t* salts,
              uint64_t salt_seed, uint32_t flags, uint8_t* proof_out, size_t proof_cap, size_t* proof_len) {
    int rc = guarded([&] {
        if (!c) throw ArgError("circuit is null");
        if (flags & PF_WITNESS_RESIDENT) throw ArgError("ZKB_WITNESS_RESIDENT is an engine flag");
        // every argument is validated BEFORE the asynchronous upload reads the caller's buffer: an early error return never
        // leaves a copy in flight (prove_resident drains the stream on every later error path)
        c->impl->validate_prove_args(public_inputs, n_pi, flags, proof_out, proof_cap);
        c->impl->upload_witness(wires, /*wait=*/false);      // queued ahead of the proof's kernels: one stream wait fewer
        size_t n = c->impl->prove_resident(public_inputs, n_pi, salts, salt_seed, flags, proof_out, proof_cap);
        if (proof_len) *proof_len = n;
        return (int)ZKB_OK;
    });
    if (rc == ZKB_E_BUFFER && proof_len && c) *proof_len = c->impl->common().proof_size();
    return rc;
}

// ---- proof engine ----
int zkb_engine_create(const uint8_t* common_bin, size_t common_len, const uint64_t* const_sigma, int is_values,
                      const uint64_t circuit_digest[4], int device, int n_contexts, int n_slots, zkb_engine** out) {
    return guarded([&] {
        if (!out) throw ArgError("out is null");
        *out = nullptr;
        if (!common_bin || !const_sigma) throw ArgError("null argument");
        (void)parse_common_data(common_bin, common_len);
        require_device(device);
        auto e = std::make_unique<zkb_engine>();
        e->impl = std::make_unique<Engine>(common_bin, common_len, const_sigma, is_values != 0, circuit_digest, device, n_contexts, n_slots);
        *out = e.release();
        return (int)ZKB_OK;
    });
}
int zkb_engine_destroy(zkb_engine* e) {
    return guarded([&] { delete e; return (int)ZKB_OK; });
}
size_t zkb_engine_proof_size(const zkb_engine* e) { return e ? e->impl->common().proof_size() : 0; }
int zkb_engine_acquire(zkb_engine* e, uint64_t** wires_buf) {
    int slot = -1;
    int rc = guarded([&] {
        if (!e || !wires_buf) throw ArgError("null argument");
        slot = e->impl->acquire(wires_buf);
        return (int)ZKB_OK;
    });
    return rc == ZKB_OK ? slot : rc;
}
int zkb_engine_release(zkb_engine* e, int slot) {
    return guarded([&] {
        if (!e) throw ArgError("engine is null");
        e->impl->release(slot);
        return (int)ZKB_OK;
    });
}
int zkb_engine_submit(zkb_engine* e, int slot, const uint64_t* public_inputs, size_t n_pi, const uint64_t* salts, uint64_t salt_seed,
                      uint32_t flags, uint8_t* proof_out, size_t proof_cap) {
    return guarded([&] {
        if (!e) throw ArgError("engine is null");
        e->impl->submit(slot, public_inputs, n_pi, salts, salt_seed, flags, proof_out, proof_cap);
        return (int)ZKB_OK;
    });
}
int zkb_engine_wait(zkb_engine* e, int slot, size_t* proof_len) {
    int rc = guarded([&] {
        if (!e) throw ArgError("engine is null");
        std::string err;
        int st = e->impl->wait(slot, proof_len, &err);
        if (st != ZKB_OK) g_last_error = err;
        return st;
    });
    return rc;
}
int zkb_last_timings(const zkb_circuit* c, float* ms_out, int cap) {
    if (!c || !ms_out) return 0;
    int n = cap < (int)T_COUNT ? cap : (int)T_COUNT;
    for (int i = 0; i < n; ++i) ms_out[i] = c->impl->timings[i];
    return n;
}

int zkb_partial_products(zkb_circuit* c, const uint64_t* wires, const uint64_t* betas, const uint64_t* gammas, uint64_t* out) {
    return guarded([&] {
        if (!c) throw ArgError("circuit is null");
        c->impl->partial_products(wires, betas, gammas, out);
        return (int)ZKB_OK;
    });
}
int zkb_quotient(zkb_circuit* c, const uint64_t* wires, const uint64_t* zs_pp, const uint64_t* public_inputs, size_t n_pi,
                 const uint64_t* betas, const uint64_t* gammas, const uint64_t* alphas, uint64_t* out) {
    return guarded([&] {
        if (!c) throw ArgError("circuit is null");
        c->impl->quotient(wires, zs_pp, public_inputs, n_pi, betas, gammas, alphas, out);
        return (int)ZKB_OK;
    });
}

// ---- standalone stage entry points ----
int zkb_poseidon_permute_batch(uint64_t* states, size_t count, int device) {
    return guarded([&] {
        if (!states && count) throw ArgError("states is null");
        if (!count) return (int)ZKB_OK;
        check_canon(states, count * 12, "states");
        require_device(device);
        DevBuf d(count * 12);
        cuda_check(cudaMemcpy(d.get(), states, count * 96, cudaMemcpyHostToDevice), "H2D");
        launch_poseidon_permute(d.get(), count, 0);
        cuda_check(cudaGetLastError(), "poseidon_permute launch");
        cuda_check(cudaMemcpy(states, d.get(), count * 96, cudaMemcpyDeviceToHost), "D2H");
        return (int)ZKB_OK;
    });
}

int zkb_lde_batch(const uint64_t* values, size_t ncols, size_t n, unsigned rate_bits, int from_coeffs, uint64_t* coeffs_out,
                  uint64_t* lde_out, int device) {
    return guarded([&] {
        if (!ncols) return (int)ZKB_OK;                      // empty batch: nothing to do
        if (!values) throw ArgError("values is null");
        if (!is_pow2(n) || n < 2 || rate_bits > 4) throw ArgError("n must be a power of two >= 2 and rate_bits <= 4");
        check_canon(values, ncols * n, "values");
        require_device(device);
        unsigned lg_n = lg2(n);
        size_t N = n << rate_bits;
        DevBuf c(ncols * n), l(lde_out ? ncols * N : 0);
        cuda_check(cudaMemcpy(c.get(), values, ncols * n * 8, cudaMemcpyHostToDevice), "H2D");
        if (!from_coeffs) launch_intt_natural(c.get(), n, c.get(), n, (int)ncols, lg_n, nullptr, 0);
        if (lde_out) launch_lde(c.get(), n, l.get(), N, (int)ncols, lg_n, rate_bits, GL_GEN, 0);
        cuda_check(cudaGetLastError(), "lde launch");
        if (coeffs_out) cuda_check(cudaMemcpy(coeffs_out, c.get(), ncols * n * 8, cudaMemcpyDeviceToHost), "D2H coeffs");
        if (lde_out) cuda_check(cudaMemcpy(lde_out, l.get(), ncols * N * 8, cudaMemcpyDeviceToHost), "D2H lde");
        cuda_check(cudaDeviceSynchronize(), "sync");
        return (int)ZKB_OK;
    });
}

int zkb_merkle_commit(const uint64_t* leaves, size_t width, size_t num_leaves, unsigned cap_height, uint64_t* digests_out,
                      uint64_t* cap_out, int device) {
    return guarded([&] {
        if (!leaves || !cap_out) throw ArgError("null argument");
        if (!is_pow2(num_leaves) || width == 0) throw ArgError("num_leaves must be a power of two and width > 0");
        if ((size_t(1) << cap_height) > num_leaves) throw ArgError("cap_height exceeds tree height");
        check_canon(leaves, width * num_leaves, "leaves");
        require_device(device);
        DevBuf d(width * num_leaves);
        size_t nd = merkle_digest_count(num_leaves, cap_height);
        DevBuf dg(nd * 4);
        cuda_check(cudaMemcpy(d.get(), leaves, width * num_leaves * 8, cudaMemcpyHostToDevice), "H2D");
        size_t cap_off = launch_merkle_tree(d.get(), num_leaves, (int)width, num_leaves, dg.get(), cap_height, 0);
        cuda_check(cudaGetLastError(), "merkle launch");
        if (digests_out) cuda_check(cudaMemcpy(digests_out, dg.get(), nd * 32, cudaMemcpyDeviceToHost), "D2H digests");
        cuda_check(cudaMemcpy(cap_out, dg.get() + cap_off * 4, (size_t(32)) << cap_height, cudaMemcpyDeviceToHost), "D2H cap");
        return (int)ZKB_OK;
    });
}

int zkb_commit_batch(const uint64_t* values, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, int reps,
                     uint64_t* cap_out, float* times_ms, int device) {
    return guarded([&] {
        if (!values || !cap_out) throw ArgError("null argument");
        if (!is_pow2(n) || n < 2 || rate_bits > 4 || !ncols) throw ArgError("bad shape");
        if (reps < 1) reps = 1;
        require_device(device);
        unsigned lg_n = lg2(n);
        size_t N = n << rate_bits;
        if ((size_t(1) << cap_height) > N) throw ArgError("cap_height exceeds tree height");
        DevBuf v(ncols * n), c(ncols * n), l(ncols * N);
        size_t nd = merkle_digest_count(N, cap_height);
        DevBuf dg(nd * 4);
        cuda_check(cudaMemcpy(v.get(), values, ncols * n * 8, cudaMemcpyHostToDevice), "H2D");
        cudaEvent_t e0, e1, e2;
        cuda_check(cudaEventCreate(&e0), "event"); cuda_check(cudaEventCreate(&e1), "event"); cuda_check(cudaEventCreate(&e2), "event");
        float t_lde = 0, t_merkle = 0;
        size_t cap_off = 0;
        // with reps > 1 one untimed pass comes first: the first transform of a size builds its twiddle / coset tables
        // (cudaMalloc + a setup launch + a sync inside launch_lde), which is not part of the steady-state time
        for (int r = reps > 1 ? -1 : 0; r < reps; ++r) {
            cuda_check(cudaEventRecord(e0, 0), "record");
            launch_intt_natural(v.get(), n, c.get(), n, (int)ncols, lg_n, nullptr, 0);
            launch_lde(c.get(), n, l.get(), N, (int)ncols, lg_n, rate_bits, GL_GEN, 0);
            cuda_check(cudaEventRecord(e1, 0), "record");
            cap_off = launch_merkle_tree(l.get(), N, (int)ncols, N, dg.get(), cap_height, 0);
            cuda_check(cudaEventRecord(e2, 0), "record");
            cuda_check(cudaEventSynchronize(e2), "sync");
            float a, b;
            cuda_check(cudaEventElapsedTime(&a, e0, e1), "elapsed");
            cuda_check(cudaEventElapsedTime(&b, e1, e2), "elapsed");
            if (r >= 0) { t_lde += a; t_merkle += b; }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        if (times_ms) { times_ms[0] = t_lde / reps; times_ms[1] = t_merkle / reps; }
        cuda_check(cudaMemcpy(cap_out, dg.get() + cap_off * 4, (size_t(32)) << cap_height, cudaMemcpyDeviceToHost), "D2H cap");
        return (int)ZKB_OK;
    });
}

int zkb_commit_cosets(const uint64_t* values, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, unsigned blk_lo,
                      unsigned blk_hi, int reps, uint64_t* cap_part_out, float* times_ms, int device) {
    return guarded([&] {
        if (!values || !cap_part_out) throw ArgError("null argument");
        if (!is_pow2(n) || n < 2 || rate_bits > 4 || !ncols) throw ArgError("bad shape");
        const unsigned nblk_all = 1u << rate_bits;
        if (blk_lo >= blk_hi || blk_hi > nblk_all || !is_pow2(blk_hi - blk_lo) || blk_lo % (blk_hi - blk_lo)) throw ArgError("bad block range");
        if (cap_height < rate_bits) throw ArgError("coset sharding needs cap_height >= rate_bits (each block must hold whole cap subtrees)");
        if (reps < 1) reps = 1;
        require_device(device);
        const unsigned lg_n = lg2(n), nb = blk_hi - blk_lo;
        const size_t N = n << rate_bits, L = n * nb;                   // local leaves
        if ((size_t(1) << cap_height) > N) throw ArgError("cap_height exceeds tree height");
        const unsigned cap_local = cap_height - rate_bits + lg2(nb);    // this rank's cap digests = 2^cap_local
        DevBuf v(ncols * n), c(ncols * n), l(ncols * L);
        DevBuf dg(merkle_digest_count(L, cap_local) * 4);
        cuda_check(cudaMemcpy(v.get(), values, ncols * n * 8, cudaMemcpyHostToDevice), "H2D");
        cudaEvent_t e0, e1, e2;
        cuda_check(cudaEventCreate(&e0), "event"); cuda_check(cudaEventCreate(&e1), "event"); cuda_check(cudaEventCreate(&e2), "event");
        float t_lde = 0, t_merkle = 0;
        size_t cap_off = 0;
        // with reps > 1 one untimed pass comes first: the first transform of a size builds its twiddle / coset tables
        // (cudaMalloc + a setup launch + a sync inside launch_lde), which is not part of the steady-state time
        for (int r = reps > 1 ? -1 : 0; r < reps; ++r) {
            cuda_check(cudaEventRecord(e0, 0), "record");
            launch_intt_natural(v.get(), n, c.get(), n, (int)ncols, lg_n, nullptr, 0);      // every rank needs all coefficients
            launch_lde_blocks(c.get(), n, l.get(), L, (int)ncols, lg_n, rate_bits, GL_GEN, blk_lo, blk_hi, 0);
            cuda_check(cudaEventRecord(e1, 0), "record");
            cap_off = launch_merkle_tree(l.get(), L, (int)ncols, L, dg.get(), cap_local, 0);
            cuda_check(cudaEventRecord(e2, 0), "record");
            cuda_check(cudaEventSynchronize(e2), "sync");
            float a, b;
            cuda_check(cudaEventElapsedTime(&a, e0, e1), "elapsed");
            cuda_check(cudaEventElapsedTime(&b, e1, e2), "elapsed");
            if (r >= 0) { t_lde += a; t_merkle += b; }
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
        if (times_ms) { times_ms[0] = t_lde / reps; times_ms[1] = t_merkle / reps; }
        cuda_check(cudaMemcpy(cap_part_out, dg.get() + cap_off * 4, (size_t(32)) << cap_local, cudaMemcpyDeviceToHost), "D2H cap");
        return (int)ZKB_OK;
    });
}

// ---- multi-GPU sharding over NCCL ----
int zkb_comm_unique_id(uint8_t id_out[ZKB_COMM_ID_BYTES]) {
    return guarded([&] {
        if (!id_out) throw ArgError("id_out is null");
        Comm::unique_id(id_out);
        return (int)ZKB_OK;
    });
}
int zkb_comm_create(const uint8_t id[ZKB_COMM_ID_BYTES], int nranks, int rank, int device, zkb_comm** out) {
    return guarded([&] {
        if (!out) throw ArgError("out is null");
        *out = nullptr;
        if (!id) throw ArgError("id is null");
        require_device(device);
        auto c = std::make_unique<zkb_comm>();
        c->impl = std::make_unique<Comm>(id, nranks, rank, device);
        *out = c.release();
        return (int)ZKB_OK;
    });
}
int zkb_comm_destroy(zkb_comm* c) {
    return guarded([&] { delete c; return (int)ZKB_OK; });
}
int zkb_commit_sharded(zkb_comm* c, const uint64_t* values, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, int reps,
                       uint64_t* cap_out, float* times_ms) {
    return guarded([&] {
        if (!c) throw ArgError("comm is null");
        c->impl->commit(values, ncols, n, rate_bits, cap_height, reps, cap_out, times_ms);
        return (int)ZKB_OK;
    });
}
int zkb_comm_peer_windows(zkb_comm* c) {
    return guarded([&] {
        if (!c) throw ArgError("comm is null");
        return c->impl->peer_windows() ? 1 : 0;
    });
}
int zkb_quotient_chunks_sharded(zkb_comm* c, const uint64_t* q_values, size_t num_challenges, size_t n, unsigned rate_bits,
                                uint64_t* chunks_out, float* times_ms) {
    return guarded([&] {
        if (!c) throw ArgError("comm is null");
        c->impl->quotient_chunks(q_values, num_challenges, n, rate_bits, chunks_out, times_ms);
        return (int)ZKB_OK;
    });
}

}  // extern "C"
