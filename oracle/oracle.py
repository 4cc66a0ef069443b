"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding over oracle/liboracle.so (the CPU restatement of the reference's proving path; see the
headers of oracle/*.hpp for the reference file:line each piece follows). Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the product
(zk-circuits_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

P = 0xFFFFFFFF00000001
u64p = ctypes.POINTER(ctypes.c_uint64)
u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile liboracle.so (idempotent)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_last_error.restype = ctypes.c_char_p
        L.orc_fmul.restype = ctypes.c_uint64
        L.orc_fmul.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.orc_finv.restype = ctypes.c_uint64
        L.orc_finv.argtypes = [ctypes.c_uint64]
        L.orc_root_of_unity.restype = ctypes.c_uint64
        L.orc_root_of_unity.argtypes = [ctypes.c_uint]
        L.orc_salt_value.restype = ctypes.c_uint64
        L.orc_salt_value.argtypes = [ctypes.c_uint64, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint64]
        for name in ("orc_common_roundtrip", "orc_proof_roundtrip", "orc_prove", "orc_trace_get", "orc_trace_challenges"):
            getattr(L, name).restype = ctypes.c_long
        for name in ("orc_synth_make", "orc_circuit_create", "orc_trace_new"):
            getattr(L, name).restype = ctypes.c_void_p
        for name in ("orc_synth_common_len", "orc_synth_degree"):
            getattr(L, name).restype = ctypes.c_size_t
        _lib = L
    return _lib


def _err():
    return lib().orc_last_error().decode()


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(u64p)


def _u8(b):
    a = np.frombuffer(bytes(b), dtype=np.uint8).copy()
    return a, a.ctypes.data_as(u8p)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def set_fast(on):
    """True: the optimised host forms of the hot loops (CPU-baseline arm of bench.py); False (default): the readable
    restatement the parity tests are written against. Same results bit for bit (tests/test_oracle_kats.py)."""
    lib().orc_set_fast(int(bool(on)))


def round_constants():
    out = np.zeros(360, dtype=np.uint64)
    lib().orc_round_constants(out.ctypes.data_as(u64p))
    return out


def poseidon_permute(states):
    """states: (count, 12) uint64 → permuted copy."""
    a = np.array(states, dtype=np.uint64, order="C").reshape(-1, 12).copy()
    lib().orc_poseidon_permute(a.ctypes.data_as(u64p), ctypes.c_size_t(a.shape[0]))
    return a


def hash_no_pad(v):
    a, p = _u64(v)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_no_pad(p, ctypes.c_size_t(a.size), out.ctypes.data_as(u64p))
    return out


def hash_pad(v):
    a, p = _u64(v)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_hash_pad(p, ctypes.c_size_t(a.size), out.ctypes.data_as(u64p))
    return out


def two_to_one(l, r):
    la, lp = _u64(l)
    ra, rp = _u64(r)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_two_to_one(lp, rp, out.ctypes.data_as(u64p))
    return out


def fmul(a, b):
    return lib().orc_fmul(a, b)


def finv(a):
    return lib().orc_finv(a)


def root_of_unity(k):
    return lib().orc_root_of_unity(k)


def salt_value(seed, batch, s, leaf):
    return lib().orc_salt_value(seed, batch, s, leaf)


def lde_batch(values, rate_bits=3, from_coeffs=False, want_lde=True):
    """values: (ncols, n) → (coeffs (ncols, n), lde (ncols, n<<rate_bits) in leaf (bit-reversed) order)."""
    a, p = _u64(values)
    ncols, n = a.shape
    coeffs = np.zeros((ncols, n), dtype=np.uint64)
    lde = np.zeros((ncols, n << rate_bits), dtype=np.uint64) if want_lde else None
    rc = lib().orc_lde_batch(p, ctypes.c_size_t(ncols), ctypes.c_size_t(n), rate_bits, int(from_coeffs),
                             coeffs.ctypes.data_as(u64p), lde.ctypes.data_as(u64p) if want_lde else None)
    if rc != 0:
        raise RuntimeError(_err())
    return coeffs, lde


def ntt(data, inverse=False):
    a = np.array(data, dtype=np.uint64, order="C").copy()
    ncols, n = a.shape
    if lib().orc_ntt(a.ctypes.data_as(u64p), ctypes.c_size_t(ncols), ctypes.c_size_t(n), int(inverse)) != 0:
        raise RuntimeError(_err())
    return a


def merkle_commit(leaves_colmajor, cap_height):
    """leaves_colmajor: (width, num_leaves). Returns (digests [all levels concatenated, (k,4)], cap (2^cap_height, 4))."""
    a, p = _u64(leaves_colmajor)
    width, nl = a.shape
    lg = nl.bit_length() - 1
    total = sum(nl >> k for k in range(lg - cap_height + 1))
    digests = np.zeros((total, 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().orc_merkle_commit(p, ctypes.c_size_t(width), ctypes.c_size_t(nl), cap_height,
                                 digests.ctypes.data_as(u64p), cap.ctypes.data_as(u64p))
    if rc != 0:
        raise RuntimeError(_err())
    return digests, cap


def verify(common, cap, digest, proof):
    """Returns "" if accepted, else the rejection reason."""
    ca, cp = _u8(common)
    pa, pp = _u8(proof)
    capa, capp = _u64(cap)
    da, dp = _u64(digest)
    rc = lib().orc_verify(cp, ctypes.c_size_t(ca.size), capp, ctypes.c_size_t(capa.size // 4), dp, pp, ctypes.c_size_t(pa.size))
    return "" if rc == 0 else (_err() or "rejected")


def verify_with_verifier_bin(vbin, proof):
    va, vp = _u8(vbin)
    pa, pp = _u8(proof)
    rc = lib().orc_verify_with_verifier_bin(vp, ctypes.c_size_t(va.size), pp, ctypes.c_size_t(pa.size))
    return "" if rc == 0 else (_err() or "rejected")


def common_roundtrip(common):
    ca, cp = _u8(common)
    out = np.zeros(ca.size + 64, dtype=np.uint8)
    n = lib().orc_common_roundtrip(cp, ctypes.c_size_t(ca.size), out.ctypes.data_as(u8p), ctypes.c_size_t(out.size))
    if n < 0:
        raise RuntimeError(_err())
    return out[:n].tobytes()


def proof_roundtrip(common, proof):
    ca, cp = _u8(common)
    pa, pp = _u8(proof)
    out = np.zeros(pa.size + 64, dtype=np.uint8)
    n = lib().orc_proof_roundtrip(cp, ctypes.c_size_t(ca.size), pp, ctypes.c_size_t(pa.size), out.ctypes.data_as(u8p), ctypes.c_size_t(out.size))
    if n < 0:
        raise RuntimeError(_err())
    return out[:n].tobytes()


_INFO_KEYS = ["degree_bits", "num_wires", "num_routed_wires", "num_constants", "num_challenges", "num_partial_products",
              "quotient_degree_factor", "zero_knowledge", "rate_bits", "cap_height", "num_query_rounds", "pow_bits",
              "num_public_inputs", "num_gates", "n_arity"]


def common_info(common):
    ca, cp = _u8(common)
    info = np.zeros(len(_INFO_KEYS), dtype=np.uint64)
    ar = np.zeros(16, dtype=np.uint64)
    if lib().orc_common_info(cp, ctypes.c_size_t(ca.size), info.ctypes.data_as(u64p), ar.ctypes.data_as(u64p)) != 0:
        raise RuntimeError(_err())
    d = {k: int(v) for k, v in zip(_INFO_KEYS, info)}
    d["reduction_arity_bits"] = [int(x) for x in ar[: d["n_arity"]]]
    return d


def challenges(vbin, proof):
    va, vp = _u8(vbin)
    pa, pp = _u8(proof)
    out = np.zeros(256, dtype=np.uint64)
    n = lib().orc_challenges(vp, ctypes.c_size_t(va.size), pp, ctypes.c_size_t(pa.size), out.ctypes.data_as(u64p), ctypes.c_size_t(out.size))
    if n < 0:
        raise RuntimeError(_err())
    v = [int(x) for x in out[:n]]
    return {"betas": v[0:2], "gammas": v[2:4], "alphas": v[4:6], "zeta": v[6:8], "fri_alpha": v[8:10],
            "pow_response": v[10], "query_indices": v[11:]}


class Synth:
    """Synthetic wormhole-/voting-shaped circuit + witness (oracle/circuit_maker.cpp)."""

    WORMHOLE = dict(n_poseidon=488, n_base_sum=3800, n_arith=2520, n_const=100, num_public_inputs=16)
    VOTING = dict(n_poseidon=34, n_base_sum=33, n_arith=120, n_const=12, num_public_inputs=13)
    TINY = dict(n_poseidon=6, n_base_sum=5, n_arith=6, n_const=3, num_public_inputs=5)

    # recursion-shaped: rows of ArithmeticExtension, MulExtension, Reducing, ReducingExtension, RandomAccess,
    # Exponentiation, CosetInterpolation, PoseidonMds on top of the base counts
    RECURSION_KEYS = ("n_arith_ext", "n_mul_ext", "n_reducing", "n_reducing_ext", "n_random_access", "n_exp", "n_coset", "n_mds")
    RECURSION = dict(n_poseidon=1600, n_base_sum=260, n_arith=500, n_const=60, num_public_inputs=16, n_arith_ext=800,
                     n_mul_ext=160, n_reducing=120, n_reducing_ext=120, n_random_access=230, n_exp=60, n_coset=112, n_mds=8)
    RECURSION_TINY = dict(n_poseidon=5, n_base_sum=3, n_arith=4, n_const=3, num_public_inputs=5, n_arith_ext=4, n_mul_ext=3,
                          n_reducing=3, n_reducing_ext=3, n_random_access=3, n_exp=3, n_coset=3, n_mds=2)

    def __init__(self, zk=False, seed=1, min_degree_bits=0, n_poseidon=488, n_base_sum=3800, n_arith=2520, n_const=100,
                 num_public_inputs=16, **recursion):
        L = lib()
        if recursion:
            unknown = set(recursion) - set(self.RECURSION_KEYS)
            if unknown:
                raise ValueError(f"bad recursion spec {unknown}")
            counts = (ctypes.c_size_t * 8)(*[int(recursion.get(k, 0)) for k in self.RECURSION_KEYS])
            L.orc_synth_make_recursion.restype = ctypes.c_void_p
            L.orc_synth_make_recursion.argtypes = [ctypes.c_uint, ctypes.c_int] + [ctypes.c_size_t] * 5 + [ctypes.c_uint64, ctypes.c_void_p]
            self._h = L.orc_synth_make_recursion(min_degree_bits, int(zk), n_poseidon, n_base_sum, n_arith, n_const, num_public_inputs, seed,
                                                 ctypes.cast(counts, ctypes.c_void_p))
        else:
            L.orc_synth_make.argtypes = [ctypes.c_uint, ctypes.c_int] + [ctypes.c_size_t] * 5 + [ctypes.c_uint64]
            self._h = L.orc_synth_make(min_degree_bits, int(zk), n_poseidon, n_base_sum, n_arith, n_const, num_public_inputs, seed)
        if not self._h:
            raise RuntimeError(_err())
        h = ctypes.c_void_p(self._h)
        n = L.orc_synth_degree(h)
        clen = L.orc_synth_common_len(h)
        cb = np.zeros(clen, dtype=np.uint8)
        L.orc_synth_common(h, cb.ctypes.data_as(u8p))
        self.common = cb.tobytes()
        self.info = common_info(self.common)
        self.n = n
        self.const_sigma_values = np.zeros((self.info["num_constants"] + self.info["num_routed_wires"], n), dtype=np.uint64)
        L.orc_synth_const_sigma_values(h, self.const_sigma_values.ctypes.data_as(u64p))
        self.wires = np.zeros((self.info["num_wires"], n), dtype=np.uint64)
        L.orc_synth_wires(h, self.wires.ctypes.data_as(u64p))
        self.public_inputs = np.zeros(num_public_inputs, dtype=np.uint64)
        L.orc_synth_public_inputs(h, self.public_inputs.ctypes.data_as(u64p))

    def check(self):
        rc = lib().orc_synth_check(ctypes.c_void_p(self._h))
        return "" if rc == 0 else _err()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_synth_free(ctypes.c_void_p(self._h))
            self._h = None


class Circuit:
    """Oracle circuit context: CommonCircuitData + constants/sigmas (values over H) → commitment + digest."""

    def __init__(self, common, const_sigma_values):
        ca, cp = _u8(common)
        va, vp = _u64(const_sigma_values)
        self._h = lib().orc_circuit_create(cp, ctypes.c_size_t(ca.size), vp)
        if not self._h:
            raise RuntimeError(_err())
        self.common = bytes(common)
        self.info = common_info(common)
        self.n = 1 << self.info["degree_bits"]
        h = ctypes.c_void_p(self._h)
        self.cap = np.zeros((1 << self.info["cap_height"], 4), dtype=np.uint64)
        lib().orc_circuit_cap(h, self.cap.ctypes.data_as(u64p))
        self.digest = np.zeros(4, dtype=np.uint64)
        lib().orc_circuit_digest(h, self.digest.ctypes.data_as(u64p))

    def const_sigma_coeffs(self):
        out = np.zeros((self.info["num_constants"] + self.info["num_routed_wires"], self.n), dtype=np.uint64)
        lib().orc_circuit_const_sigma_coeffs(ctypes.c_void_p(self._h), out.ctypes.data_as(u64p))
        return out

    def prove(self, wires, public_inputs, salts=None, salt_seed=0, trace=False):
        wa, wp = _u64(wires)
        pa, pp = _u64(public_inputs)
        sp = None
        if salts is not None:
            sa, sp = _u64(salts)
        out = np.zeros(1 << 20, dtype=np.uint8)
        tr = ctypes.c_void_p(lib().orc_trace_new()) if trace else None
        n = lib().orc_prove(ctypes.c_void_p(self._h), wp, pp, ctypes.c_size_t(pa.size), sp, ctypes.c_uint64(salt_seed),
                            out.ctypes.data_as(u8p), ctypes.c_size_t(out.size), tr)
        if n < 0:
            if tr:
                lib().orc_trace_free(tr)
            raise RuntimeError(_err())
        proof = out[:n].tobytes()
        if not trace:
            return proof
        t = Trace(tr, self.info)
        return proof, t

    def partial_products(self, wires, betas, gammas):
        wa, wp = _u64(wires)
        ba, bp = _u64(betas)
        ga, gp = _u64(gammas)
        k = self.info["num_challenges"] * (1 + self.info["num_partial_products"])
        out = np.zeros((k, self.n), dtype=np.uint64)
        if lib().orc_partial_products(ctypes.c_void_p(self._h), wp, bp, gp, out.ctypes.data_as(u64p)) != 0:
            raise RuntimeError(_err())
        return out

    def quotient(self, wires, zs_pp, public_inputs, betas, gammas, alphas):
        wa, wp = _u64(wires)
        za, zp = _u64(zs_pp)
        pa, pp = _u64(public_inputs)
        ba, bp = _u64(betas)
        ga, gp = _u64(gammas)
        aa, ap = _u64(alphas)
        k = self.info["num_challenges"] * self.info["quotient_degree_factor"]
        out = np.zeros((k, self.n), dtype=np.uint64)
        if lib().orc_quotient(ctypes.c_void_p(self._h), wp, zp, pp, ctypes.c_size_t(pa.size), bp, gp, ap, out.ctypes.data_as(u64p)) != 0:
            raise RuntimeError(_err())
        return out

    def verify(self, proof):
        return verify(self.common, self.cap, self.digest, proof)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_circuit_free(ctypes.c_void_p(self._h))
            self._h = None


class Trace:
    def __init__(self, handle, info):
        self._h = handle
        self.info = info
        out = np.zeros(64, dtype=np.uint64)
        n = lib().orc_trace_challenges(handle, out.ctypes.data_as(u64p), ctypes.c_size_t(out.size))
        v = [int(x) for x in out[:n]]
        self.betas, self.gammas, self.alphas = v[0:2], v[2:4], v[4:6]
        self.zeta, self.fri_alpha = v[6:8], v[8:10]
        self.fri_betas = [v[10 + 2 * i: 12 + 2 * i] for i in range((n - 10) // 2)]

    def get(self, which):
        n = lib().orc_trace_get(self._h, which, None, ctypes.c_size_t(0))
        if n < 0:
            raise KeyError(which)
        out = np.zeros(n, dtype=np.uint64)
        lib().orc_trace_get(self._h, which, out.ctypes.data_as(u64p), ctypes.c_size_t(n))
        return out

    @property
    def zs_pp_values(self):
        return self.get(0).reshape(-1, 1 << self.info["degree_bits"])

    @property
    def quotient_chunks(self):
        return self.get(1).reshape(-1, 1 << self.info["degree_bits"])

    @property
    def final_poly_pre_fri(self):
        return self.get(2).reshape(-1, 2)

    def fri_layer_values(self, i):
        return self.get(3 + i).reshape(-1, 2)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_trace_free(self._h)
            self._h = None
