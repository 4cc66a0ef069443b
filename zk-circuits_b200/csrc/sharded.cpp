#include "sharded.hpp"
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace zkb {

namespace {
// the NCCL entry points used here, resolved at run time
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
const NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        const char* names[] = {std::getenv("ZKB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        void* h = nullptr;
        for (const char* nm : names)
            if (nm && (h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
        if (!h) { err = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror(); return; }
        auto sym = [&](const char* s) { void* p = dlsym(h, s); if (!p && err.empty()) err = std::string("NCCL symbol missing: ") + s; return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    if (!err.empty()) throw NcclError(err);
    return api;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw NcclError(std::string(what) + ": " + nccl().GetErrorString(r));
}
#define NK(x) nccl_check((x), #x)
#define CK(x) cuda_check((x), #x)
unsigned lg2u(size_t x) { unsigned k = 0; while ((size_t(1) << k) < x) ++k; return k; }
struct Events {
    std::vector<cudaEvent_t> ev;
    explicit Events(size_t n, bool timing = true) : ev(n, nullptr) {
        for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
    }
    ~Events() { for (auto e : ev) if (e) cudaEventDestroy(e); }
    cudaEvent_t operator[](size_t i) const { return ev[i]; }
};
}  // namespace

static_assert(sizeof(ncclUniqueId) == COMM_ID_BYTES, "ncclUniqueId size");

void Comm::unique_id(uint8_t out[COMM_ID_BYTES]) {
    ncclUniqueId id;
    NK(nccl().GetUniqueId(&id));
    std::memcpy(out, &id, COMM_ID_BYTES);
}

Comm::Comm(const uint8_t id_bytes[COMM_ID_BYTES], int nranks, int rank, int device) : nranks_(nranks), rank_(rank), device_(device) {
    if (nranks < 1 || rank < 0 || rank >= nranks) throw ArgError("bad rank / nranks");
    if (nranks & (nranks - 1)) throw ArgError("the number of ranks must be a power of two");
    CK(cudaSetDevice(device));
    device_tables_init(device);
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, COMM_ID_BYTES);
    ncclComm_t c = nullptr;
    NK(nccl().CommInitRank(&c, nranks, id, rank));
    comm_ = c;
    try {
        CK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&comm_st_, cudaStreamNonBlocking));
    } catch (...) {
        if (st_) cudaStreamDestroy(st_);
        nccl().CommDestroy(c);
        throw;
    }
}
// ---- peer windows: every rank's coefficient buffer mapped into every other rank (CUDA IPC over NVLink) ----
void Comm::barrier() {
    if (!sync_words_) CK(cudaMalloc(&sync_words_, sizeof(u64) * (size_t)nranks_));
    NK(nccl().AllGather(sync_words_ + rank_, sync_words_, 1, ncclUint64, static_cast<ncclComm_t>(comm_), st_));
}
void Comm::release_window() {
    if (!win_) return;
    for (int r = 0; r < nranks_; ++r)
        if (r != rank_ && r < (int)peer_win_.size() && peer_win_[r]) cudaIpcCloseMemHandle(peer_win_[r]);
    peer_win_.clear();
    try {                                   // nobody may still have this window mapped when it is freed
        barrier();
        cudaStreamSynchronize(st_);
    } catch (...) {}
    cudaFree(win_);
    win_ = nullptr;
    win_words_ = 0;
}
bool Comm::ensure_window(size_t words) {
    static const bool enabled = [] { const char* e = std::getenv("ZKB_SHARDED_P2P"); return !(e && e[0] == '0'); }();
    if (!enabled || nranks_ < 2) return false;
    if (p2p_tried_ && !p2p_ok_) return false;
    if (words <= win_words_) return p2p_ok_;
    // every rank takes the same path: the shapes of a sharded call are the same on all of them
    release_window();
    p2p_tried_ = true;
    const size_t G = (size_t)nranks_, HW = sizeof(cudaIpcMemHandle_t) / sizeof(u64) + 1;      // handle words + one flag word
    ncclComm_t comm = static_cast<ncclComm_t>(comm_);
    std::vector<u64> mine(HW, 0), all(G * HW, 0);
    cudaIpcMemHandle_t h;
    if (cudaMalloc(&win_, words * sizeof(u64)) == cudaSuccess && cudaIpcGetMemHandle(&h, win_) == cudaSuccess) {
        std::memcpy(mine.data(), &h, sizeof h);
        mine[HW - 1] = 1;
    } else {
        cudaGetLastError();
    }
    DevBuf hb(G * HW);
    CK(cudaMemcpyAsync(hb.get() + (size_t)rank_ * HW, mine.data(), HW * 8, cudaMemcpyHostToDevice, st_));
    NK(nccl().AllGather(hb.get() + (size_t)rank_ * HW, hb.get(), HW, ncclUint64, comm, st_));
    CK(cudaMemcpyAsync(all.data(), hb.get(), G * HW * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    u64 ok = 1;
    for (size_t r = 0; r < G; ++r) ok &= all[r * HW + HW - 1];
    peer_win_.assign(G, nullptr);
    if (ok) {
        for (size_t r = 0; r < G && ok; ++r) {
            if ((int)r == rank_) { peer_win_[r] = win_; continue; }
            cudaIpcMemHandle_t ph;
            std::memcpy(&ph, &all[r * HW], sizeof ph);
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
            peer_win_[r] = static_cast<u64*>(ptr);
        }
    }
    // agree: one rank that could not map a peer sends everybody to the NCCL gather
    if (!sync_words_) CK(cudaMalloc(&sync_words_, sizeof(u64) * G));
    CK(cudaMemcpyAsync(sync_words_ + rank_, &ok, 8, cudaMemcpyHostToDevice, st_));
    NK(nccl().AllGather(sync_words_ + rank_, sync_words_, 1, ncclUint64, comm, st_));
    std::vector<u64> oks(G);
    CK(cudaMemcpyAsync(oks.data(), sync_words_, G * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    for (u64 v : oks) ok &= v;
    p2p_ok_ = ok != 0;
    if (p2p_ok_) win_words_ = words;
    else release_window();
    return p2p_ok_;
}

Comm::~Comm() {
    cudaSetDevice(device_);
    release_window();
    if (sync_words_) cudaFree(sync_words_);
    if (st_) { cudaStreamSynchronize(st_); cudaStreamDestroy(st_); }
    if (comm_st_) { cudaStreamSynchronize(comm_st_); cudaStreamDestroy(comm_st_); }
    if (comm_) nccl().CommDestroy(static_cast<ncclComm_t>(comm_));
}

void Comm::commit(const u64* values, size_t ncols, size_t n, unsigned rate_bits, unsigned cap_height, int reps, u64* cap_out,
                  float* times_ms) {
    if (!values || !cap_out) throw ArgError("null argument");
    if (n < 2 || (n & (n - 1)) || rate_bits > 4 || !ncols) throw ArgError("bad shape");
    const unsigned G = (unsigned)nranks_, R = 1u << rate_bits;
    if (G > R) throw ArgError("more ranks than coset blocks");
    if (cap_height < rate_bits) throw ArgError("coset sharding needs cap_height >= rate_bits (each block must hold whole cap subtrees)");
    if (reps < 1) reps = 1;
    CK(cudaSetDevice(device_));
    ncclComm_t comm = static_cast<ncclComm_t>(comm_);
    const unsigned lg_n = lg2u(n), B = R / G, blk_lo = (unsigned)rank_ * B;
    const size_t N = n << rate_bits, L = n * B;
    if ((size_t(1) << cap_height) > N) throw ArgError("cap_height exceeds tree height");
    const unsigned cap_local = cap_height - rate_bits + lg2u(B);
    // columns in chunks of G * w: rank r interpolates columns [k G w + r w, + w) of chunk k (zero padding past ncols)
    const size_t per_rank = (ncols + G - 1) / G;
    const bool p2p = ensure_window(per_rank * n);          // peers pull my coefficients out of my window: one chunk per rank
    const size_t nchunks = (!p2p && per_rank >= 8) ? 4 : 1, w = (per_rank + nchunks - 1) / nchunks, padded = nchunks * G * w;
    DevBuf coeffs(padded * n), lde(ncols * L), dg(merkle_digest_count(L, cap_local) * 4), cap_all((size_t(4) << cap_height));
    DevBuf vals(nchunks * w * n);
    CK(cudaMemsetAsync(vals.get(), 0, nchunks * w * n * 8, st_));
    for (size_t k = 0; k < nchunks; ++k) {      // H2D of this rank's column slices only: 1 / G of the matrix
        const size_t c0 = k * G * w + (size_t)rank_ * w;
        if (c0 >= ncols) continue;
        const size_t cnt = std::min(w, ncols - c0);
        CK(cudaMemcpyAsync(vals.get() + k * w * n, values + c0 * n, cnt * n * 8, cudaMemcpyHostToDevice, st_));
    }
    Events tm(5), chain(std::max<size_t>(2 * nchunks, G), false);
    float t_lde = 0, t_merkle = 0, t_gather = 0;
    size_t cap_off = 0;
    auto pass_p2p = [&] {
        CK(cudaEventRecord(tm[0], st_));
        launch_intt_natural(vals.get(), n, win_, n, (int)w, lg_n, nullptr, st_);
        CK(cudaEventRecord(tm[3], st_));
        barrier();                                         // every window holds its owner's coefficients
        CK(cudaEventRecord(chain[rank_], st_));
        CK(cudaStreamWaitEvent(comm_st_, chain[rank_], 0));
        // pull the other ranks' slices out of their windows with the copy engines (NVLink), nearest rank first so that no window
        // is read by every peer at once; the LDE of a slice starts when its copy has landed, the rank's own slice needs none
        for (unsigned s1 = 1; s1 < G; ++s1) {
            const size_t r = ((size_t)rank_ + s1) % G, c0 = r * w;
            if (c0 >= ncols) continue;
            const size_t cnt = std::min(w, ncols - c0);
            CK(cudaMemcpyAsync(coeffs.get() + c0 * n, peer_win_[r], cnt * n * 8, cudaMemcpyDefault, comm_st_));
            CK(cudaEventRecord(chain[r], comm_st_));
        }
        CK(cudaEventRecord(tm[4], comm_st_));
        for (unsigned s1 = 0; s1 < G; ++s1) {
            const size_t r = ((size_t)rank_ + s1) % G, c0 = r * w;
            if (c0 >= ncols) continue;
            const size_t cnt = std::min(w, ncols - c0);
            if (s1) CK(cudaStreamWaitEvent(st_, chain[r], 0));
            launch_lde_blocks(s1 ? coeffs.get() + c0 * n : win_, n, lde.get() + c0 * L, L, (int)cnt, lg_n, rate_bits, GL_GEN, blk_lo,
                              blk_lo + B, st_);
        }
        CK(cudaEventRecord(tm[1], st_));
        cap_off = launch_merkle_tree(lde.get(), L, (int)ncols, L, dg.get(), cap_local, st_);
        // the cap gather is also the barrier after which a window may be overwritten: a rank contributes once its copies have run
        NK(nccl().AllGather(dg.get() + cap_off * 4, cap_all.get(), (size_t(4) << cap_local), ncclUint64, comm, st_));
        CK(cudaEventRecord(tm[2], st_));
        CK(cudaEventSynchronize(tm[2]));
        CK(cudaStreamSynchronize(comm_st_));
    };
    auto pass_gather = [&] {
        CK(cudaEventRecord(tm[0], st_));
        for (size_t k = 0; k < nchunks; ++k) {             // interpolate my slice of chunk k, then gather the chunk (comm stream)
            u64* mine = coeffs.get() + (k * G * w + (size_t)rank_ * w) * n;
            launch_intt_natural(vals.get() + k * w * n, n, mine, n, (int)w, lg_n, nullptr, st_);
            CK(cudaEventRecord(chain[2 * k], st_));
            CK(cudaStreamWaitEvent(comm_st_, chain[2 * k], 0));
            if (k == 0) CK(cudaEventRecord(tm[3], comm_st_));
            NK(nccl().AllGather(mine, coeffs.get() + k * G * w * n, w * n, ncclUint64, comm, comm_st_));   // in place
            CK(cudaEventRecord(chain[2 * k + 1], comm_st_));
        }
        CK(cudaEventRecord(tm[4], comm_st_));
        for (size_t k = 0; k < nchunks; ++k) {             // the gather of chunk k + 1 runs behind the LDE of chunk k
            CK(cudaStreamWaitEvent(st_, chain[2 * k + 1], 0));
            const size_t c0 = k * G * w;
            if (c0 >= ncols) break;
            const size_t cnt = std::min(G * w, ncols - c0);
            launch_lde_blocks(coeffs.get() + c0 * n, n, lde.get() + c0 * L, L, (int)cnt, lg_n, rate_bits, GL_GEN, blk_lo, blk_lo + B, st_);
        }
        CK(cudaEventRecord(tm[1], st_));
        cap_off = launch_merkle_tree(lde.get(), L, (int)ncols, L, dg.get(), cap_local, st_);
        NK(nccl().AllGather(dg.get() + cap_off * 4, cap_all.get(), (size_t(4) << cap_local), ncclUint64, comm, st_));
        CK(cudaEventRecord(tm[2], st_));
        CK(cudaEventSynchronize(tm[2]));
        CK(cudaStreamSynchronize(comm_st_));
    };
    auto pass = [&] { if (p2p) pass_p2p(); else pass_gather(); };
    if (reps > 1) pass();                                  // untimed: twiddle / coset tables, NCCL channel setup
    for (int r = 0; r < reps; ++r) {
        pass();
        float a, b, c;
        CK(cudaEventElapsedTime(&a, tm[0], tm[1]));
        CK(cudaEventElapsedTime(&b, tm[1], tm[2]));
        CK(cudaEventElapsedTime(&c, tm[3], tm[4]));
        t_lde += a; t_merkle += b; t_gather += c;
    }
    if (times_ms) { times_ms[0] = t_lde / reps; times_ms[1] = t_merkle / reps; times_ms[2] = t_gather / reps; }
    CK(cudaMemcpyAsync(cap_out, cap_all.get(), (size_t(32)) << cap_height, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
}

void Comm::quotient_chunks(const u64* q_values, size_t nch, size_t n, unsigned rate_bits, u64* chunks_out, float* times_ms) {
    if (!q_values || !chunks_out) throw ArgError("null argument");
    if (n < 2 || (n & (n - 1)) || rate_bits > 4 || !nch || nch > 4) throw ArgError("bad shape");
    const unsigned G = (unsigned)nranks_, R = 1u << rate_bits;
    if (G > R) throw ArgError("more ranks than coset blocks");
    if (n % G) throw ArgError("n must be a multiple of the number of ranks");
    CK(cudaSetDevice(device_));
    ncclComm_t comm = static_cast<ncclComm_t>(comm_);
    const unsigned lg_n = lg2u(n), B = R / G;
    const size_t sl = n / G;                                  // coefficient indices per rank
    const u64 w_N = gl_root_of_unity(lg_n + rate_bits);
    const bool p2p = ensure_window(nch * B * n);              // the interpolants live in the window the peers have mapped
    DevBuf u_local(p2p ? 1 : nch * B * n), all((size_t)nch * R * sl), out((size_t)nch * R * sl), mat((size_t)R * R);
    u64* const u = p2p ? win_ : u_local.get();
    Events tm(3);
    CK(cudaMemcpyAsync(u, q_values, nch * B * n * 8, cudaMemcpyHostToDevice, st_));
    // V^-1[m][j] = c_0^-m w_R^(-j m) / R with c_j = (g w_N^j)^n = c_0 w_R^j
    std::vector<u64> vinv((size_t)R * R);
    const u64 c0_inv = gl_inv(gl_pow(GL_GEN, n)), wR_inv = gl_inv(gl_pow(w_N, n)), r_inv = gl_inv(R);
    for (unsigned m = 0; m < R; ++m)
        for (unsigned j = 0; j < R; ++j)
            vinv[(size_t)m * R + j] = gl_mul(gl_mul(gl_pow(c0_inv, m), gl_pow(wR_inv, (u64)j * m)), r_inv);
    CK(cudaMemcpyAsync(mat.get(), vinv.data(), vinv.size() * 8, cudaMemcpyHostToDevice, st_));
    CK(cudaEventRecord(tm[0], st_));
    // u_j: interpolant of t on coset j (block i of this rank is leaf block jb = rank B + i, coset j = bitrev(jb))
    for (size_t ch = 0; ch < nch; ++ch)
        for (unsigned i = 0; i < B; ++i) {
            const unsigned jb = (unsigned)rank_ * B + i, j = bitrev32(jb, rate_bits);
            launch_coset_intt_bitrev(u + (ch * B + i) * n, n, 1, lg_n, gl_mul(GL_GEN, gl_pow(w_N, j)), st_);
        }
    CK(cudaEventRecord(tm[1], st_));
    // all-to-all: this rank needs coefficient slice `rank` of every (challenge, block) interpolant of every peer, placed by coset
    if (p2p) {                      // pulled out of the peers' windows over NVLink (copy engines), nearest rank first
        barrier();                  // every window holds its interpolants
        for (unsigned s1 = 0; s1 < G; ++s1) {
            const unsigned p = ((unsigned)rank_ + s1) % G;
            for (size_t ch = 0; ch < nch; ++ch)
                for (unsigned i = 0; i < B; ++i) {
                    const unsigned j = bitrev32(p * B + i, rate_bits);
                    CK(cudaMemcpyAsync(all.get() + (ch * R + j) * sl, peer_win_[p] + (ch * B + i) * n + (size_t)rank_ * sl, sl * 8,
                                       cudaMemcpyDefault, st_));
                }
        }
        barrier();                  // nobody overwrites its window while a peer still reads it
    } else {
        NK(nccl().GroupStart());
        for (unsigned p = 0; p < G; ++p)
            for (size_t ch = 0; ch < nch; ++ch)
                for (unsigned i = 0; i < B; ++i) {
                    NK(nccl().Send(u + (ch * B + i) * n + (size_t)p * sl, sl, ncclUint64, (int)p, comm, st_));
                    const unsigned j = bitrev32(p * B + i, rate_bits);
                    NK(nccl().Recv(all.get() + (ch * R + j) * sl, sl, ncclUint64, (int)p, comm, st_));
                }
        NK(nccl().GroupEnd());
    }
    for (size_t ch = 0; ch < nch; ++ch)
        launch_vandermonde_solve(all.get() + ch * R * sl, out.get() + ch * R * sl, sl, R, mat.get(), st_);
    CK(cudaEventRecord(tm[2], st_));
    CK(cudaMemcpyAsync(chunks_out, out.get(), (size_t)nch * R * sl * 8, cudaMemcpyDeviceToHost, st_));
    CK(cudaStreamSynchronize(st_));
    if (times_ms) {
        CK(cudaEventElapsedTime(&times_ms[0], tm[0], tm[1]));
        CK(cudaEventElapsedTime(&times_ms[1], tm[1], tm[2]));
    }
}

}  // namespace zkb
