"""The C-ABI shared library loads and exports every symbol include/zkb200.h declares. CPU only:
no compute call is made unless a GPU is present (there is no CPU fallback to call)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def zkb():
    import zkb200

    zkb200.build()
    return zkb200


def declared_symbols(header="zkb200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(zkb):
    syms = declared_symbols()
    assert len(syms) >= 15
    L = ctypes.CDLL(zkb.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/zkb200.h but not exported"
    assert sorted(zkb.EXPORTS) == syms
    assert "sm_100a" in zkb.version()
    # the synthetic workload generator is test / bench tooling: its own header and library, nothing of it in libzkb200.so
    ssyms = declared_symbols("zkb200_synth.h")
    S = ctypes.CDLL(zkb.SYNTH_LIB_PATH)
    for s in ssyms:
        assert hasattr(S, s) and not hasattr(L, s), s
    assert sorted(zkb.SYNTH_EXPORTS) == ssyms


def test_status_codes_match_header(zkb):
    src = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    for code, name in zkb.STATUS.items():
        m = re.search(name + r"\s*=\s*(-?\d+)", src)
        assert m and int(m.group(1)) == code


def test_malformed_common_data_is_a_parse_error(zkb):
    # parse errors are reported before any device is touched
    from conftest import golden_bytes

    good = golden_bytes("bench_common.bin")
    cs = np.zeros(84 * 16, dtype=np.uint64)
    for bad in (good[:-1], good + b"\0", good[:100], b""):
        with pytest.raises(zkb.ZkbError) as e:
            zkb.ProverCircuit(bad if bad else b"\0", cs)
        assert e.value.status == "ZKB_E_PARSE"
    # unsupported gate tag (6 = Lookup; lookup tables are outside the implemented set) -> ZKB_E_UNSUPPORTED_GATE
    tampered = bytearray(good)
    tampered[997] = 6
    with pytest.raises(zkb.ZkbError) as e:
        zkb.ProverCircuit(bytes(tampered), cs)
    assert e.value.status == "ZKB_E_UNSUPPORTED_GATE"


def test_no_cpu_fallback(zkb):
    if zkb.device_count() > 0:
        pytest.skip("a GPU is present; the fail-loudly path is exercised on the CPU box")
    with pytest.raises(zkb.ZkbError) as e:
        zkb.poseidon_permute_batch(np.zeros((1, 12), dtype=np.uint64))
    assert e.value.status == "ZKB_E_CUDA"
    with pytest.raises(zkb.ZkbError) as e:
        zkb.lde_batch(np.zeros((1, 8), dtype=np.uint64))
    assert e.value.status == "ZKB_E_CUDA"
    # the round-2 entry points fail the same way: no engine, no communicator, no sharded commitment without a device
    from conftest import golden_bytes

    cs = np.zeros((84, 1 << 14), dtype=np.uint64)
    with pytest.raises(zkb.ZkbError) as e:
        zkb.Engine(golden_bytes("bench_common.bin"), cs, contexts=2)
    assert e.value.status == "ZKB_E_CUDA"
    with pytest.raises(zkb.ZkbError) as e:
        zkb.Comm(np.zeros(128, dtype=np.uint8), 1, 0)
    assert e.value.status == "ZKB_E_CUDA"


def test_flag_values_match_header(zkb):
    src = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    for name, val in (("ZKB_POW_MIN", zkb.POW_MIN), ("ZKB_SALTS_FROM_SEED", zkb.SALTS_FROM_SEED), ("ZKB_CHECK_WITNESS", zkb.CHECK_WITNESS),
                      ("ZKB_WITNESS_RESIDENT", zkb.WITNESS_RESIDENT)):
        m = re.search(r"#define " + name + r"\s+(0x[0-9a-fA-F]+|\d+)u", src)
        assert m and int(m.group(1), 0) == val, name
    rs = open(os.path.join(ROOT, "rust", "zkb200-sys", "src", "lib.rs")).read()
    for name, val in (("ZKB_SALTS_FROM_SEED", 0x100), ("ZKB_CHECK_WITNESS", 0x200), ("ZKB_WITNESS_RESIDENT", 0x400)):
        assert re.search(r"pub const " + name + r": u32 = " + hex(val) + ";", rs), name


def test_rust_sys_crate_declares_every_header_symbol():
    """rust/zkb200-sys cannot be compiled here (no Rust toolchain); keep it at least in step with the header."""
    rs = open(os.path.join(ROOT, "rust", "zkb200-sys", "src", "lib.rs")).read()
    declared = set(re.findall(r"pub fn (zkb_[a-z_0-9]+)\s*\(", rs))
    assert declared == set(declared_symbols())
    hdr = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    for name, val in re.findall(r"pub const (ZKB_E_[A-Z_]+): c_int = (-?\d+);", rs):
        assert re.search(name + r"\s*=\s*" + val + r"\b", hdr), name


def test_header_is_plain_c_and_the_c_example_links(zkb, tmp_path):
    """include/zkb200.h must be usable from C (no C++ types in the ABI): compile and link examples/prove_example.c with
    gcc -std=c11 against the built library. Without a GPU the program exits 2 after printing the library version."""
    import subprocess

    exe = tmp_path / "prove_example"
    lib_dir = os.path.dirname(zkb.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-O2", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "prove_example.c"), "-L", lib_dir, "-lzkb200", "-lzkb200_synth",
                           f"-Wl,-rpath,{lib_dir}", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    if zkb.device_count() == 0:
        assert out.returncode == 2 and "no CUDA device" in out.stderr and "sm_100a" in out.stderr
    else:
        assert out.returncode == 0, out.stderr


@pytest.mark.parametrize("shape", ["tiny", "tiny_zk", "recursion", "recursion_zk"])
def test_workload_generator_matches_the_oracles(zkb, oracle, shape):
    """zkb_synth_* (host code standing in for the Rust circuit builder + witness generation) emits the same circuit,
    witness and CommonCircuitData bytes as the oracle's generator for the same spec — for the wormhole gate set and for the
    recursion gate set — and the oracle's checker accepts the witness."""
    if shape.startswith("recursion"):
        spec = dict(seed=3, zk=shape.endswith("_zk"), **oracle.Synth.RECURSION_TINY)
    else:
        spec = dict(seed=6, zk=shape == "tiny_zk", **oracle.Synth.TINY)
    a = zkb.SynthCircuit(**spec)
    b = oracle.Synth(**spec)
    assert b.check() == ""
    assert a.common == b.common
    assert np.array_equal(a.const_sigma_values, b.const_sigma_values)
    assert np.array_equal(a.wires, b.wires)
    assert np.array_equal(a.public_inputs, b.public_inputs)
