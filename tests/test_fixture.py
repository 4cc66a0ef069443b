"""L4 fixture directory format (rust/zkb200::Fixture::write_dir <-> zkb200.fixture): CPU round trip with the oracle standing in
for the Rust host. The GPU replay of the same directories is tests/test_gpu_new_paths.py."""
import numpy as np
import pytest


@pytest.mark.parametrize("zk", [False, True])
def test_fixture_directory_round_trip(oracle, tmp_path, zk):
    from zkb200 import fixture

    s = oracle.Synth(zk=zk, seed=9, **oracle.Synth.TINY)
    oc = oracle.Circuit(s.common, s.const_sigma_values)
    salts = None
    if zk:
        salts = np.random.default_rng(5).integers(0, oracle.P, size=(3, 4, s.n * 8), dtype=np.uint64)
    proof = oc.prove(s.wires, s.public_inputs, salts=salts)
    fixture.write(tmp_path / "fx", s.common, oc.const_sigma_coeffs(), oc.digest, s.wires, s.public_inputs, proof, salts=salts)
    fx = fixture.load(tmp_path / "fx")
    assert fx["common"] == s.common and not fx["is_values"] and fx["proof"] == proof
    assert np.array_equal(fx["const_sigma"], oc.const_sigma_coeffs())
    assert np.array_equal(fx["wires"], s.wires) and np.array_equal(fx["public_inputs"], s.public_inputs)
    assert np.array_equal(fx["circuit_digest"], oc.digest)
    assert (fx["salts"] is None) == (not zk) and (not zk or np.array_equal(fx["salts"], salts))
    assert oc.verify(fx["proof"]) == ""


def test_rust_exporter_writes_the_same_file_names():
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rs = open(os.path.join(root, "rust", "zkb200", "src", "lib.rs")).read()
    names = set(re.findall(r'dir\.join\("([^"]+)"\)', rs))
    assert names == {"common.bin", "const_sigma_coeffs.u64", "circuit_digest.u64", "wires.u64", "public_inputs.u64", "salts.u64", "proof.bin"}
    assert "pub salts: Option<Vec<u64>>" in rs
