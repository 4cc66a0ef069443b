"""L4 byte-parity fixtures: a directory written by `rust/zkb200::Fixture::write_dir` on a Rust host —
the reference's own circuit data, witness, salts and CPU proof — replayed through `zkb_prove`.

Layout (raw little-endian u64 arrays unless noted):
    common.bin               CommonCircuitData::to_bytes            (circuit-builder/src/lib.rs:36-39)
    const_sigma_coeffs.u64   [(num_constants + num_routed)][n], coefficient form  (or const_sigma_values.u64: values over H)
    circuit_digest.u64       4 words
    wires.u64                [num_wires][n]
    public_inputs.u64
    salts.u64                [3][4][8n], zero-knowledge circuits only
    proof.bin                ProofWithPublicInputs::to_bytes of the CPU prover (RAYON_NUM_THREADS=1)

The same writer is used by the tests here with the CPU oracle standing in for the Rust host, so the loader and the replay path
(explicit salts, coefficient-form constants, digest check) are exercised without a Rust toolchain.
"""
import os

import numpy as np


def _read_u64(path):
    return np.fromfile(path, dtype="<u8").astype(np.uint64)


def write(path, common, const_sigma, circuit_digest, wires, public_inputs, proof, salts=None, is_values=False):
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "common.bin"), "wb") as f:
        f.write(bytes(common))
    name = "const_sigma_values.u64" if is_values else "const_sigma_coeffs.u64"
    np.ascontiguousarray(const_sigma, dtype="<u8").tofile(os.path.join(path, name))
    np.ascontiguousarray(circuit_digest, dtype="<u8").tofile(os.path.join(path, "circuit_digest.u64"))
    np.ascontiguousarray(wires, dtype="<u8").tofile(os.path.join(path, "wires.u64"))
    np.ascontiguousarray(public_inputs, dtype="<u8").tofile(os.path.join(path, "public_inputs.u64"))
    if salts is not None:
        np.ascontiguousarray(salts, dtype="<u8").tofile(os.path.join(path, "salts.u64"))
    with open(os.path.join(path, "proof.bin"), "wb") as f:
        f.write(bytes(proof))


def load(path):
    """Returns a dict with common (bytes), const_sigma [cols][n], is_values, circuit_digest, wires [num_wires][n],
    public_inputs, salts (or None), proof (bytes). Shapes are derived from the header of common.bin."""
    with open(os.path.join(path, "common.bin"), "rb") as f:
        common = f.read()
    head = np.frombuffer(common[:24], dtype="<u8")
    num_wires, num_routed = int(head[0]), int(head[1])
    values_path = os.path.join(path, "const_sigma_values.u64")
    is_values = os.path.exists(values_path)
    cs = _read_u64(values_path if is_values else os.path.join(path, "const_sigma_coeffs.u64"))
    wires = _read_u64(os.path.join(path, "wires.u64"))
    if wires.size % num_wires:
        raise ValueError("wires.u64 is not a multiple of num_wires")
    n = wires.size // num_wires
    if n & (n - 1) or cs.size % n or cs.size // n <= num_routed:
        raise ValueError("fixture shapes are inconsistent with common.bin")
    salts_path = os.path.join(path, "salts.u64")
    salts = _read_u64(salts_path) if os.path.exists(salts_path) else None
    if salts is not None:
        if salts.size % 12:
            raise ValueError("salts.u64 must hold 3 batches x 4 columns")
        salts = salts.reshape(3, 4, salts.size // 12)
    with open(os.path.join(path, "proof.bin"), "rb") as f:
        proof = f.read()
    return dict(common=common, const_sigma=cs.reshape(-1, n), is_values=is_values,
                circuit_digest=_read_u64(os.path.join(path, "circuit_digest.u64")), wires=wires.reshape(num_wires, n),
                public_inputs=_read_u64(os.path.join(path, "public_inputs.u64")), salts=salts, proof=proof)


def replay(path, device=0, check_witness=True):
    """Prove the fixture's witness with the CUDA prover; returns (gpu_proof_bytes, fixture_proof_bytes)."""
    from . import ProverCircuit

    fx = load(path)
    circ = ProverCircuit(fx["common"], fx["const_sigma"], is_values=fx["is_values"], circuit_digest=fx["circuit_digest"], device=device)
    got = circ.prove(fx["wires"], fx["public_inputs"], salts=fx["salts"], check_witness=check_witness)
    return got, fx["proof"]
