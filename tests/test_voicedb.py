"""voice.db format: our writer/parser against the reference's layout (ctts.h:84-111, ctts.c:1000-1080)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest


def test_small_db_is_the_golden_one(H, golden, small_db):
    # the synthetic voice is seeded; golden vectors were made with exactly these bytes
    assert hashlib.sha256(small_db).digest() == golden["db_sha256"].tobytes()


def test_header_and_index_layout(H, small_db):
    db = H.voicedb.parse_voice_db(small_db)
    assert db.unit_count == 38 + 1100
    assert db.index.dtype.itemsize == 32
    # units sorted by char_count desc then byte order (compare_units, ctts.c:931)
    keys = [(-int(e["char_count"]), db.unit_text(i).encode()) for i, e in enumerate(db.index)]
    assert keys == sorted(keys)
    # audio offsets are a running sum in samples
    cnt = db.index["sample_count"].astype(np.int64)
    assert np.array_equal(db.index["audio_offset"].astype(np.int64), np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    assert int(cnt.sum()) == db.total_samples
    assert db.hash_table_size & (db.hash_table_size - 1) == 0 and db.hash_table_size >= db.unit_count / 0.7


def test_hash_chains_find_every_unit(H, small_db):
    db = H.voicedb.parse_voice_db(small_db)
    for i in range(0, db.unit_count, 7):
        t = db.unit_text(i).encode()
        h = H.voicedb.fnv1a(t)
        j = int(db.hash_table[h % db.hash_table_size])
        while j != 0xFFFFFFFF and j != i:
            j = int(db.index[j]["next_hash"])
        assert j == i


def test_builder_is_byte_identical_to_reference_build(H, tmp_path):
    if not H.have_reference():
        pytest.skip("oracle/_ref not built")
    lu, su = H.voicedb.synthetic_units(60, seed=5)
    lu, su = lu[:12], su[:60]
    ours = H.voicedb.build_voice_db(lu + su)
    H.voicedb.write_dataset(str(tmp_path), lu, su)
    out = tmp_path / "ref.db"
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        rc = H.ref_lib().ref_build_database(str(tmp_path).encode(), str(out).encode())
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    assert rc == 0
    assert out.read_bytes() == ours


def test_rejects_bad_magic(H, small_db):
    bad = b"XXXX" + small_db[4:]
    with pytest.raises(ValueError):
        H.voicedb.parse_voice_db(bad)
    with pytest.raises(RuntimeError):
        H.front.Front(bad, H.shipped_config(), None)
