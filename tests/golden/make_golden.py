#!/usr/bin/env python3
"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container only;
/root/reference does not exist on the GPU box, which is why the outputs are committed).

  * binary fixtures the reference's own benches/tests load:
      wormhole/bench-data/{common,verifier,proof}.bin   (verifier/benches/verifier.rs:11,16-25)
      wormhole/aggregator/data/dummy_proof{,_zk}.bin    (aggregator/src/util.rs:6-9)
  * kats.json: the known-answer vectors embedded in the reference's Rust tests, extracted by regex:
      wormhole/tests/src/circuit/unspendable_account_tests.rs:12-27   (5 secret -> address pairs)
      wormhole/tests/src/prover/prover_tests.rs:31-41                 (nullifier, root hash bytes)
      wormhole/tests/test-helpers/src/lib.rs:10-23,65-81              (default secret/accounts, storage proof)
"""
import hashlib
import json
import os
import re
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

BINS = {
    "bench_common.bin": "wormhole/bench-data/common.bin",
    "bench_verifier.bin": "wormhole/bench-data/verifier.bin",
    "bench_proof.bin": "wormhole/bench-data/proof.bin",
    "dummy_proof.bin": "wormhole/aggregator/data/dummy_proof.bin",
    "dummy_proof_zk.bin": "wormhole/aggregator/data/dummy_proof_zk.bin",
}


def rust_str_array(src, name):
    m = re.search(name + r"[^=]*=\s*\[(.*?)\];", src, re.S)
    return re.findall(r'"([0-9a-fA-F]+)"', m.group(1))


def rust_u8_array(src, anchor):
    m = re.search(anchor + r".*?\[(.*?)\]", src, re.S)
    return [int(x) for x in re.findall(r"\d+", m.group(1))]


def main():
    sha = {}
    for dst, src in BINS.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        sha[dst] = hashlib.sha256(open(os.path.join(HERE, dst), "rb").read()).hexdigest()

    ua = open(os.path.join(REF, "wormhole/tests/src/circuit/unspendable_account_tests.rs")).read()
    pt = open(os.path.join(REF, "wormhole/tests/src/prover/prover_tests.rs")).read()
    th = open(os.path.join(REF, "wormhole/tests/test-helpers/src/lib.rs")).read()
    kats = {
        "sources": {
            "secrets/addresses": "wormhole/tests/src/circuit/unspendable_account_tests.rs:12-27",
            "nullifier/root_hash": "wormhole/tests/src/prover/prover_tests.rs:31-41",
            "defaults/storage_proof": "wormhole/tests/test-helpers/src/lib.rs:10-23,65-81",
        },
        "secrets": rust_str_array(ua, "SECRETS"),
        "addresses": rust_str_array(ua, "ADDRESSES"),
        "nullifier": rust_u8_array(pt, r"nullifier: BytesDigest::try_from\("),
        "root_hash_bytes": rust_u8_array(pt, r"root_hash: BytesDigest::try_from\("),
        "default_secret": re.search(r'DEFAULT_SECRET: &str = "([0-9a-f]+)"', th).group(1),
        "default_transfer_count": int(re.search(r"DEFAULT_TRANSFER_COUNT: u64 = (\d+)", th).group(1)),
        "default_to_account": rust_u8_array(th, r"DEFAULT_TO_ACCOUNT: \[u8; 32\] ="),
        "default_funding_account": rust_u8_array(th, r"DEFAULT_FUNDING_ACCOUNT: \[u8; 32\] ="),
        "default_root_hash": re.search(r'DEFAULT_ROOT_HASH: &str =\s*"([0-9a-f]+)"', th).group(1),
        "storage_proof": rust_str_array(th, "DEFAULT_STORAGE_PROOF: "),
        "storage_proof_indices": [int(x) for x in re.findall(r"\d+", re.search(r"DEFAULT_STORAGE_PROOF_INDICIES[^=]*=\s*\[(.*?)\]", th, re.S).group(1))],
        "unspendable_salt": "wormhole",
        "nullifier_salt": "~nullif~",
        "sha256": sha,
    }
    assert len(kats["secrets"]) == 5 and len(kats["addresses"]) == 5
    assert len(kats["nullifier"]) == 32 and len(kats["root_hash_bytes"]) == 32
    assert len(kats["storage_proof"]) == 7 and len(kats["storage_proof_indices"]) == 7
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", sorted(BINS), "and kats.json")


if __name__ == "__main__":
    main()
