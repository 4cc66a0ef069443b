"""zkb200 — ctypes binding over libzkb200.so, the B200-native Plonky2 proving backend.

Host-side mirror of the reference's prove boundary for tests and benchmarks:
`ProverCircuit(common_bin, const_sigma, ...)` plays `ProverCircuitData` (prover_only + common,
/root/reference/wormhole/prover/src/lib.rs:114-130) and `ProverCircuit.prove(wires, public_inputs)` plays
`circuit_data.prove(partial_witness)` (/root/reference/wormhole/prover/src/lib.rs:233-237) after witness
generation, returning `ProofWithPublicInputs::to_bytes()`.

There is no CPU fallback: loading fails loudly if the CUDA library has not been built, and every call
fails with ZkbError(ZKB_E_CUDA) when no sm_100 device is present.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("ZKB200_LIB") or os.path.join(_ROOT, "libzkb200.so")      # ZKB200_LIB: a variant build (lab experiments)

u64p = ctypes.POINTER(ctypes.c_uint64)
u8p = ctypes.POINTER(ctypes.c_uint8)
f32p = ctypes.POINTER(ctypes.c_float)

STATUS = {0: "ZKB_OK", -1: "ZKB_E_ARG", -2: "ZKB_E_PARSE", -3: "ZKB_E_UNSUPPORTED_GATE", -4: "ZKB_E_UNSAT",
          -5: "ZKB_E_ZETA_IN_SUBGROUP", -6: "ZKB_E_CUDA", -7: "ZKB_E_NCCL", -8: "ZKB_E_BUFFER", -9: "ZKB_E_DIGEST"}
TIMING_KEYS = ["wires_intt", "wires_lde", "wires_merkle", "partial_products", "zs_commit", "quotient", "quotient_commit",
               "openings", "fri_combine", "fri_commit", "pow", "queries", "total", "host_transcript", "host_permutations"]

EXPORTS = ["zkb_version", "zkb_last_error", "zkb_device_count", "zkb_kernel_launch_count", "zkb_circuit_create", "zkb_circuit_destroy",
           "zkb_circuit_verifier_only", "zkb_proof_size", "zkb_prove", "zkb_witness_upload", "zkb_prove_resident",
           "zkb_last_timings", "zkb_poseidon_permute_batch", "zkb_lde_batch", "zkb_merkle_commit", "zkb_commit_batch",
           "zkb_commit_cosets",
           "zkb_partial_products", "zkb_quotient", "zkb_engine_create", "zkb_engine_destroy", "zkb_engine_proof_size",
           "zkb_engine_acquire", "zkb_engine_release", "zkb_engine_submit", "zkb_engine_wait", "zkb_comm_unique_id", "zkb_comm_create",
           "zkb_comm_destroy", "zkb_commit_sharded", "zkb_comm_peer_windows", "zkb_quotient_chunks_sharded"]
# `flags` of the prove calls (include/zkb200.h)
POW_MIN, SALTS_FROM_SEED, CHECK_WITNESS, WITNESS_RESIDENT = 0, 0x100, 0x200, 0x400
SYNTH_LIB_PATH = os.path.join(_ROOT, "libzkb200_synth.so")
SYNTH_EXPORTS = ["zkb_synth_last_error", "zkb_synth_create", "zkb_synth_create_recursion", "zkb_synth_num_constants",
                 "zkb_synth_destroy", "zkb_synth_common_len", "zkb_synth_degree", "zkb_synth_get"]


class ZkbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code
        self.status = STATUS.get(code, str(code))


def build(force=False):
    """Compile libzkb200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH) or not os.path.exists(SYNTH_LIB_PATH):
        subprocess.check_call(["make", "-s", "-j4", "-C", _ROOT])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {_ROOT}` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.zkb_version.restype = ctypes.c_char_p
        L.zkb_last_error.restype = ctypes.c_char_p
        L.zkb_kernel_launch_count.restype = ctypes.c_ulonglong
        L.zkb_proof_size.restype = ctypes.c_size_t
        L.zkb_proof_size.argtypes = [ctypes.c_void_p]
        L.zkb_circuit_create.argtypes = [u8p, ctypes.c_size_t, u64p, ctypes.c_int, u64p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.zkb_circuit_destroy.argtypes = [ctypes.c_void_p]
        L.zkb_circuit_verifier_only.argtypes = [ctypes.c_void_p, u64p, ctypes.c_size_t, u64p]
        L.zkb_prove.argtypes = [ctypes.c_void_p, u64p, u64p, ctypes.c_size_t, u64p, ctypes.c_uint64, ctypes.c_uint32, u8p,
                                ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.zkb_witness_upload.argtypes = [ctypes.c_void_p, u64p]
        L.zkb_prove_resident.argtypes = [ctypes.c_void_p, u64p, ctypes.c_size_t, u64p, ctypes.c_uint64, ctypes.c_uint32, u8p,
                                         ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        L.zkb_last_timings.argtypes = [ctypes.c_void_p, f32p, ctypes.c_int]
        L.zkb_poseidon_permute_batch.argtypes = [u64p, ctypes.c_size_t, ctypes.c_int]
        L.zkb_lde_batch.argtypes = [u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int, u64p, u64p, ctypes.c_int]
        L.zkb_merkle_commit.argtypes = [u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint, u64p, u64p, ctypes.c_int]
        L.zkb_commit_batch.argtypes = [u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint, ctypes.c_uint, ctypes.c_int, u64p, f32p, ctypes.c_int]
        L.zkb_commit_cosets.argtypes = [u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint,
                                        ctypes.c_int, u64p, f32p, ctypes.c_int]
        L.zkb_partial_products.argtypes = [ctypes.c_void_p, u64p, u64p, u64p, u64p]
        L.zkb_quotient.argtypes = [ctypes.c_void_p, u64p, u64p, u64p, ctypes.c_size_t, u64p, u64p, u64p, u64p]
        L.zkb_engine_create.argtypes = [u8p, ctypes.c_size_t, u64p, ctypes.c_int, u64p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_void_p)]
        L.zkb_engine_destroy.argtypes = [ctypes.c_void_p]
        L.zkb_engine_proof_size.argtypes = [ctypes.c_void_p]
        L.zkb_engine_proof_size.restype = ctypes.c_size_t
        L.zkb_engine_acquire.argtypes = [ctypes.c_void_p, ctypes.POINTER(u64p)]
        L.zkb_engine_release.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.zkb_engine_submit.argtypes = [ctypes.c_void_p, ctypes.c_int, u64p, ctypes.c_size_t, u64p, ctypes.c_uint64, ctypes.c_uint32, u8p,
                                        ctypes.c_size_t]
        L.zkb_engine_wait.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]
        L.zkb_comm_unique_id.argtypes = [u8p]
        L.zkb_comm_create.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.zkb_comm_destroy.argtypes = [ctypes.c_void_p]
        L.zkb_commit_sharded.argtypes = [ctypes.c_void_p, u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint, ctypes.c_uint, ctypes.c_int,
                                         u64p, f32p]
        L.zkb_comm_peer_windows.argtypes = [ctypes.c_void_p]
        L.zkb_quotient_chunks_sharded.argtypes = [ctypes.c_void_p, u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint, u64p, f32p]
        _lib = L
    return _lib


_synth_lib = None


def synth_lib():
    """libzkb200_synth.so: the synthetic workload generator (test / bench tooling, include/zkb200_synth.h)."""
    global _synth_lib
    if _synth_lib is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise ImportError(f"{SYNTH_LIB_PATH} is missing: build it with `make -C {_ROOT}`")
        L = ctypes.CDLL(SYNTH_LIB_PATH)
        L.zkb_synth_last_error.restype = ctypes.c_char_p
        L.zkb_synth_create_recursion.argtypes = [ctypes.c_uint, ctypes.c_int] + [ctypes.c_size_t] * 5 + [ctypes.c_uint64, ctypes.c_void_p,
                                                 ctypes.POINTER(ctypes.c_void_p)]
        L.zkb_synth_num_constants.restype = ctypes.c_size_t
        L.zkb_synth_num_constants.argtypes = [ctypes.c_void_p]
        L.zkb_synth_create.argtypes = [ctypes.c_uint, ctypes.c_int] + [ctypes.c_size_t] * 5 + [ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]
        L.zkb_synth_destroy.argtypes = [ctypes.c_void_p]
        L.zkb_synth_common_len.argtypes = [ctypes.c_void_p]
        L.zkb_synth_common_len.restype = ctypes.c_size_t
        L.zkb_synth_degree.argtypes = [ctypes.c_void_p]
        L.zkb_synth_degree.restype = ctypes.c_size_t
        L.zkb_synth_get.argtypes = [ctypes.c_void_p, u8p, u64p, u64p, u64p]
        _synth_lib = L
    return _synth_lib


def _check_synth(rc):
    if rc != 0:
        raise ZkbError(-1, synth_lib().zkb_synth_last_error().decode())


def _check(rc):
    if rc != 0:
        raise ZkbError(rc, lib().zkb_last_error().decode())


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(u64p)


def _ptr(addr_or_array):
    """Accept a numpy array or a raw host address (e.g. a pinned torch tensor's data_ptr())."""
    if isinstance(addr_or_array, int):
        return None, ctypes.cast(addr_or_array, u64p)
    return _u64(addr_or_array)


def version():
    return lib().zkb_version().decode()


def device_count():
    return lib().zkb_device_count()


def kernel_launch_count():
    return int(lib().zkb_kernel_launch_count())


def poseidon_permute_batch(states, device=0):
    a = np.array(states, dtype=np.uint64, order="C").reshape(-1, 12).copy()
    _check(lib().zkb_poseidon_permute_batch(a.ctypes.data_as(u64p), a.shape[0], device))
    return a


def lde_batch(values, rate_bits=3, from_coeffs=False, want_lde=True, device=0):
    a, p = _u64(values)
    if a.ndim != 2:
        raise ValueError("values must be (ncols, n)")
    ncols, n = a.shape
    coeffs = np.zeros((ncols, n), dtype=np.uint64)
    lde = np.zeros((ncols, n << rate_bits), dtype=np.uint64) if want_lde else None
    _check(lib().zkb_lde_batch(p, ncols, n, rate_bits, int(from_coeffs), coeffs.ctypes.data_as(u64p),
                               lde.ctypes.data_as(u64p) if want_lde else None, device))
    return coeffs, lde


def merkle_commit(leaves_colmajor, cap_height, want_digests=True, device=0):
    a, p = _u64(leaves_colmajor)
    width, nl = a.shape
    lg = nl.bit_length() - 1
    total = sum(nl >> k for k in range(lg - cap_height + 1))
    digests = np.zeros((total, 4), dtype=np.uint64) if want_digests else None
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    _check(lib().zkb_merkle_commit(p, width, nl, cap_height, digests.ctypes.data_as(u64p) if want_digests else None,
                                   cap.ctypes.data_as(u64p), device))
    return digests, cap


def commit_batch(values, rate_bits=3, cap_height=4, reps=1, device=0):
    """Fused from_values (iNTT + LDE + Merkle). Returns (cap, {'lde_ms', 'merkle_ms'})."""
    a, p = _u64(values)
    ncols, n = a.shape
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    t = np.zeros(2, dtype=np.float32)
    _check(lib().zkb_commit_batch(p, ncols, n, rate_bits, cap_height, reps, cap.ctypes.data_as(u64p), t.ctypes.data_as(f32p), device))
    return cap, {"lde_ms": float(t[0]), "merkle_ms": float(t[1])}


def commit_cosets(values, rate_bits, cap_height, blk_lo, blk_hi, reps=1, device=0):
    """One rank's share of a coset-sharded commitment: cap digests of leaf blocks [blk_lo, blk_hi) (see zkb200.sharded)."""
    a, p = _u64(values)
    ncols, n = a.shape
    rows = max(blk_hi - blk_lo, 1) << max(cap_height - rate_bits, 0)     # bad ranges are rejected by the library
    part = np.zeros((rows, 4), dtype=np.uint64)
    t = np.zeros(2, dtype=np.float32)
    _check(lib().zkb_commit_cosets(p, ncols, n, rate_bits, cap_height, blk_lo, blk_hi, reps, part.ctypes.data_as(u64p),
                                   t.ctypes.data_as(f32p), device))
    return part, {"lde_ms": float(t[0]), "merkle_ms": float(t[1])}


def _flags(salt_seed, check_witness):
    return POW_MIN | (SALTS_FROM_SEED if salt_seed is not None else 0) | (CHECK_WITNESS if check_witness else 0)


class ProverCircuit:
    """Device-resident circuit context (constants/sigmas commitment, twiddles, work buffers)."""

    def __init__(self, common_bin, const_sigma, is_values=False, circuit_digest=None, device=0):
        self._h = ctypes.c_void_p()
        cb = np.frombuffer(bytes(common_bin), dtype=np.uint8).copy()
        cs, csp = _u64(const_sigma)
        dg = dgp = None
        if circuit_digest is not None:
            dg, dgp = _u64(circuit_digest)
        _check(lib().zkb_circuit_create(cb.ctypes.data_as(u8p), cb.size, csp, int(is_values), dgp, device, ctypes.byref(self._h)))
        self.proof_size = lib().zkb_proof_size(self._h)
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().zkb_circuit_destroy(self._h)
            self._h = None

    __del__ = close

    def verifier_only(self, cap_height=4):
        cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
        digest = np.zeros(4, dtype=np.uint64)
        _check(lib().zkb_circuit_verifier_only(self._h, cap.ctypes.data_as(u64p), cap.size, digest.ctypes.data_as(u64p)))
        return cap, digest

    def prove(self, wires, public_inputs, salts=None, salt_seed=None, check_witness=False):
        """wires: (num_wires, n) array or pinned host address; returns proof bytes. Salts of a zk circuit: explicit `salts`,
        else the device CSPRNG (default), else — salt_seed given — the deterministic test stream (ZKB_SALTS_FROM_SEED).
        check_witness: ZKB_CHECK_WITNESS (an unsatisfied witness raises ZkbError ZKB_E_UNSAT)."""
        pow_rule = _flags(salt_seed, check_witness)
        salt_seed = salt_seed or 0
        _, wp = _ptr(wires)
        pa, pp = _u64(public_inputs)
        sa = sp = None
        if salts is not None:
            sa, sp = _u64(salts)
        out = np.zeros(self.proof_size, dtype=np.uint8)
        n = ctypes.c_size_t(0)
        _check(lib().zkb_prove(self._h, wp, pp, pa.size, sp, salt_seed, pow_rule, out.ctypes.data_as(u8p), out.size, ctypes.byref(n)))
        return out[: n.value].tobytes()

    def upload_witness(self, wires):
        _, wp = _ptr(wires)
        _check(lib().zkb_witness_upload(self._h, wp))

    def prove_resident(self, public_inputs, salts=None, salt_seed=None, check_witness=False, out=None):
        pow_rule = _flags(salt_seed, check_witness)
        salt_seed = salt_seed or 0
        pa, pp = _u64(public_inputs)
        sa = sp = None
        if salts is not None:
            sa, sp = _u64(salts)
        if out is None:
            out = np.zeros(self.proof_size, dtype=np.uint8)
        n = ctypes.c_size_t(0)
        _check(lib().zkb_prove_resident(self._h, pp, pa.size, sp, salt_seed, pow_rule, out.ctypes.data_as(u8p), out.size, ctypes.byref(n)))
        return out[: n.value]

    def timings(self):
        t = np.zeros(len(TIMING_KEYS), dtype=np.float32)
        k = lib().zkb_last_timings(self._h, t.ctypes.data_as(f32p), t.size)
        return {key: float(v) for key, v in zip(TIMING_KEYS[:k], t[:k])}

    def partial_products(self, wires, betas, gammas, num_cols, n):
        wa, wp = _u64(wires)
        ba, bp = _u64(betas)
        ga, gp = _u64(gammas)
        out = np.zeros((num_cols, n), dtype=np.uint64)
        _check(lib().zkb_partial_products(self._h, wp, bp, gp, out.ctypes.data_as(u64p)))
        return out

    def quotient(self, wires, zs_pp, public_inputs, betas, gammas, alphas, num_chunks, n):
        wa, wp = _u64(wires)
        za, zp = _u64(zs_pp)
        pa, pp = _u64(public_inputs)
        ba, bp = _u64(betas)
        ga, gp = _u64(gammas)
        aa, ap = _u64(alphas)
        out = np.zeros((num_chunks, n), dtype=np.uint64)
        _check(lib().zkb_quotient(self._h, wp, zp, pp, pa.size, bp, gp, ap, out.ctypes.data_as(u64p)))
        return out


def comm_unique_id():
    """128 bytes from ncclGetUniqueId: rank 0 creates them, every rank passes them to Comm()."""
    out = np.zeros(128, dtype=np.uint8)
    _check(lib().zkb_comm_unique_id(out.ctypes.data_as(u8p)))
    return out


class Comm:
    """zkb_comm: this process's rank in an NCCL communicator owned by libzkb200.so (one process per GPU)."""

    def __init__(self, unique_id, nranks, rank, device=0):
        self._h = ctypes.c_void_p()
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        _check(lib().zkb_comm_create(uid.ctypes.data_as(u8p), nranks, rank, device, ctypes.byref(self._h)))
        self.nranks, self.rank = nranks, rank

    def close(self):
        if getattr(self, "_h", None):
            lib().zkb_comm_destroy(self._h)
            self._h = None

    __del__ = close

    def commit(self, values, rate_bits=3, cap_height=4, reps=1):
        """Coset-sharded from_values commitment; returns (cap [2^cap_height][4], {'lde_ms','merkle_ms','gather_ms'})."""
        a, p = _u64(values)
        ncols, n = a.shape
        cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
        t = np.zeros(3, dtype=np.float32)
        _check(lib().zkb_commit_sharded(self._h, p, ncols, n, rate_bits, cap_height, reps, cap.ctypes.data_as(u64p), t.ctypes.data_as(f32p)))
        return cap, {"lde_ms": float(t[0]), "merkle_ms": float(t[1]), "gather_ms": float(t[2]),
                     "peer_windows": bool(lib().zkb_comm_peer_windows(self._h) == 1)}

    def quotient_chunks(self, q_values, n, rate_bits=3):
        """q_values [nch][B n]: the quotient's evaluations on this rank's leaf blocks -> [nch][2^rate_bits][n / nranks]."""
        a, p = _u64(q_values)
        nch = a.shape[0]
        out = np.zeros((nch, 1 << rate_bits, n // self.nranks), dtype=np.uint64)
        t = np.zeros(2, dtype=np.float32)
        _check(lib().zkb_quotient_chunks_sharded(self._h, p, nch, n, rate_bits, out.ctypes.data_as(u64p), t.ctypes.data_as(f32p)))
        return out, {"interpolate_ms": float(t[0]), "exchange_ms": float(t[1])}


class Engine:
    """zkb_engine: n_contexts prover contexts of one circuit on one GPU behind an asynchronous submit / wait interface with
    pinned witness slots (include/zkb200.h). `acquire()` returns (slot, wires) where wires is a writable numpy view
    [num_wires, n] of the slot's pinned buffer."""

    def __init__(self, common_bin, const_sigma, is_values=False, circuit_digest=None, device=0, contexts=8, slots=0, num_wires=135):
        self._h = ctypes.c_void_p()
        cb = np.frombuffer(bytes(common_bin), dtype=np.uint8).copy()
        cs, csp = _u64(const_sigma)
        dg = dgp = None
        if circuit_digest is not None:
            dg, dgp = _u64(circuit_digest)
        _check(lib().zkb_engine_create(cb.ctypes.data_as(u8p), cb.size, csp, int(is_values), dgp, device, contexts, slots or 2 * contexts,
                                       ctypes.byref(self._h)))
        self.proof_size = lib().zkb_engine_proof_size(self._h)
        self.shape = (num_wires, cs.shape[-1])
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None):
            lib().zkb_engine_destroy(self._h)
            self._h = None

    __del__ = close

    def acquire(self):
        buf = u64p()
        slot = lib().zkb_engine_acquire(self._h, ctypes.byref(buf))
        if slot < 0:
            _check(slot)
        n_words = self.shape[0] * self.shape[1]
        arr = np.ctypeslib.as_array(buf, shape=(n_words,)).reshape(self.shape)
        return slot, arr

    def release(self, slot):
        _check(lib().zkb_engine_release(self._h, slot))

    def submit(self, slot, public_inputs, salts=None, salt_seed=None, check_witness=False, resident=False, out=None):
        flags = _flags(salt_seed, check_witness) | (WITNESS_RESIDENT if resident else 0)
        pa, pp = _u64(public_inputs)
        sa = sp = None
        if salts is not None:
            sa, sp = _u64(salts)
        if out is None:
            out = np.zeros(self.proof_size, dtype=np.uint8)
        _check(lib().zkb_engine_submit(self._h, slot, pp, pa.size, sp, salt_seed or 0, flags, out.ctypes.data_as(u8p), out.size))
        self._keep[slot] = (out, sa)

    def wait(self, slot):
        n = ctypes.c_size_t(0)
        # take our buffers out of the table BEFORE the C call: zkb_engine_wait frees the slot, and another thread may acquire
        # and submit on the same slot id before this thread runs again
        out, salts_keep = self._keep.pop(slot, (None, None))
        rc = lib().zkb_engine_wait(self._h, slot, ctypes.byref(n))
        del salts_keep
        _check(rc)
        return out[: n.value]

    def prove(self, wires, public_inputs, **kw):
        """Blocking convenience: acquire, copy the witness in, submit, wait."""
        slot, buf = self.acquire()
        buf[:] = wires
        self.submit(slot, public_inputs, **kw)
        return self.wait(slot).tobytes()


# row mixes of the reference circuits
WORMHOLE = dict(n_poseidon=488, n_base_sum=3800, n_arith=2520, n_const=100, num_public_inputs=16)
VOTING = dict(n_poseidon=34, n_base_sum=33, n_arith=120, n_const=12, num_public_inputs=13)
TINY = dict(n_poseidon=6, n_base_sum=5, n_arith=6, n_const=3, num_public_inputs=5)


class SynthCircuit:
    """Synthetic wormhole-/voting-shaped circuit + satisfying witness (csrc/synth.cpp; host code)."""

    RECURSION_KEYS = ("n_arith_ext", "n_mul_ext", "n_reducing", "n_reducing_ext", "n_random_access", "n_exp", "n_coset", "n_mds")
    # one aggregation chunk's recursive-verifier circuit: 2^12 rows before blinding, n = 2^14 with zk=True as the aggregator
    # configures it (row mix: an estimate)
    RECURSION = dict(n_poseidon=1600, n_base_sum=260, n_arith=500, n_const=60, num_public_inputs=16, n_arith_ext=800,
                     n_mul_ext=160, n_reducing=120, n_reducing_ext=120, n_random_access=230, n_exp=60, n_coset=112, n_mds=8)

    def __init__(self, zk=False, seed=1, min_degree_bits=0, n_poseidon=488, n_base_sum=3800, n_arith=2520, n_const=100,
                 num_public_inputs=16, **recursion):
        h = ctypes.c_void_p()
        if recursion:
            if set(recursion) - set(self.RECURSION_KEYS):
                raise ValueError("bad recursion spec")
            rows = (ctypes.c_size_t * 8)(*[int(recursion.get(k, 0)) for k in self.RECURSION_KEYS])
            _check_synth(synth_lib().zkb_synth_create_recursion(min_degree_bits, int(zk), n_poseidon, n_base_sum, n_arith, n_const, num_public_inputs,
                                                    seed, ctypes.cast(rows, ctypes.c_void_p), ctypes.byref(h)))
        else:
            _check_synth(synth_lib().zkb_synth_create(min_degree_bits, int(zk), n_poseidon, n_base_sum, n_arith, n_const, num_public_inputs,
                                          seed, ctypes.byref(h)))
        try:
            n = synth_lib().zkb_synth_degree(h)
            cb = np.zeros(synth_lib().zkb_synth_common_len(h), dtype=np.uint8)
            self.n = n
            self.zk = bool(zk)
            self.const_sigma_values = np.zeros((synth_lib().zkb_synth_num_constants(h) + 80, n), dtype=np.uint64)
            self.wires = np.zeros((135, n), dtype=np.uint64)
            self.public_inputs = np.zeros(num_public_inputs, dtype=np.uint64)
            _check_synth(synth_lib().zkb_synth_get(h, cb.ctypes.data_as(u8p), self.const_sigma_values.ctypes.data_as(u64p),
                                       self.wires.ctypes.data_as(u64p), self.public_inputs.ctypes.data_as(u64p)))
            self.common = cb.tobytes()
        finally:
            synth_lib().zkb_synth_destroy(h)
