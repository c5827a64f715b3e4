"""Generates tests/golden/ref_corpus.json from the COMPILED, UNMODIFIED reference (oracle/_ref/libctts_ref.so):
SHA-1 and length of the reference's PCM for 48 sentences of the benchmark corpus on the small test voice
(32 at speed 1.0, 16 at other speeds), and for 6 sentences on a voice whose PCM pool starts at an ODD byte
offset in voice.db (audio_offset = 64 + 32 N + 4 H + strings, arbitrary parity: ctts.c:1001-1004, :1159).

Run where the reference tree exists:   python tests/golden/make_golden_corpus.py
Samples whose reference value depends on an out-of-bounds heap read (apply_smooth_pitch_contour,
ctts.c:2243-2252; the oracle reports the spans) are zeroed before hashing, by the generator and by the tests.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402


masked_sha1 = H.masked_sha1


def corpus_cases():
    """A pool of candidates: run() keeps the first 32 usable ones at speed 1.0 and the first 16 at other speeds."""
    texts = H.corpus.batch(160, seed=20261018, target_chars=110)
    speeds = [1.0] * 64 + [float(s) for s in H.corpus.mixed_speeds(96, seed=8)]
    return texts, speeds


odd_offset_voice = H.odd_offset_voice


ODD_TEXTS = ["olá mundo", "Bom dia, como vai você?", "A casa azul fica a 12 km da praia!", "não sei", "o rato roeu a roupa",
             "rato rua rio", "um dois três", "olá mundo", "casa", "a ideia"]
ODD_SPEEDS = [1.0, 1.0, 1.0, 1.0, 1.0, 1.5, 0.7, 1.5, 2.0, 0.6]


def run(db: bytes, texts, speeds, want_plain: int, want_stretched: int):
    """Reference PCM of the first want_plain usable utterances at speed 1.0 and want_stretched at other speeds.
    Usable: the spans tainted by the reference's out-of-bounds read are all recorded (<= 8) -- and, for a
    stretched utterance, there is none (the mask is in pre-stretch coordinates)."""
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    orc = H.Oracle(db)
    plan = fr.plan(texts, speeds)
    rows = []
    with tempfile.TemporaryDirectory() as d:
        dbp = os.path.join(d, "voice.db")
        with open(dbp, "wb") as f:
            f.write(db)
        ref = H.Reference(dbp)
        for u, (t, s) in enumerate(zip(texts, speeds)):
            plain = s == 1.0
            if (plain and want_plain == 0) or (not plain and want_stretched == 0):
                continue
            got, st = orc.synth(prm, plan.utt_ops(u), s)
            if st.ub_spans > H.MAX_UB_SPANS or (not plain and st.ub_spans):
                continue
            want = ref.synth(t, s)
            mask = H.ub_mask(st, max(len(want), len(got)))
            assert len(want) == len(got), (u, len(want), len(got))
            assert masked_sha1(want, mask) == masked_sha1(got, mask), f"oracle != reference for {t!r} at {s}"
            rows.append({"text": t, "speed": s, "samples": int(len(want)), "sha1": masked_sha1(want, mask),
                         "masked": int(mask[:len(want)].sum())})
            if plain:
                want_plain -= 1
            else:
                want_stretched -= 1
    assert want_plain == 0 and want_stretched == 0, (want_plain, want_stretched)
    return rows


def main() -> None:
    assert H.have_reference(), "build oracle/_ref first (make -C oracle ref)"
    texts, speeds = corpus_cases()
    out = {"voice_sha256": hashlib.sha256(H.small_db()).hexdigest(), "corpus": run(H.small_db(), texts, speeds, 32, 16)}
    odd = odd_offset_voice()
    out["odd_voice_sha256"] = hashlib.sha256(odd).hexdigest()
    out["odd_voice"] = run(odd, ODD_TEXTS, ODD_SPEEDS, 4, 2)
    path = os.path.join(HERE, "ref_corpus.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=0)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out["corpus"]), "+", len(out["odd_voice"]), "utterances;",
          sum(r["masked"] for r in out["corpus"]), "masked samples")


if __name__ == "__main__":
    main()
