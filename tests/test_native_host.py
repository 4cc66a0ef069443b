"""Host (g++) build of the device arithmetic headers: the limb-form Poseidon in csrc/poseidon.cuh compiles for the host
too, so its algebra (FFT-style MDS, limb normalisation, lazy folds) is checked here without a GPU against the plain
u128 permutation. The GPU parity tests then only have to catch code-generation differences."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_limb_poseidon_matches_plain_permutation(tmp_path):
    exe = tmp_path / "test_poseidon_limbs"
    # -DZKB_CHECK_BOUNDS: abort if a limb ever leaves the range the 32-bit wrap-around argument assumes
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-DZKB_CHECK_BOUNDS", "-o", str(exe),
                           os.path.join(ROOT, "tests", "native", "test_poseidon_limbs.cpp")])
    out = subprocess.run([str(exe), "5000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("ok")


def test_avx2_transcript_permutation_matches_the_scalar_form(tmp_path):
    """csrc/host_poseidon_avx2.cpp (the Fiat-Shamir permutation the prover runs on AVX2 hosts) against a naive 128-bit
    restatement, on corner and random states plus the permutation KAT."""
    exe, obj = tmp_path / "test_avx2", tmp_path / "avx2.o"
    subprocess.check_call(["/usr/bin/g++", "-O3", "-std=c++17", "-mavx2", "-c", "-o", str(obj),
                           os.path.join(ROOT, "zk-circuits_b200", "csrc", "host_poseidon_avx2.cpp")])
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-o", str(exe),
                           os.path.join(ROOT, "tests", "native", "test_host_transcript_avx2.cpp"), str(obj)])
    out = subprocess.run([str(exe), "3000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("ok")
