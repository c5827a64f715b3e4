"""Host front end (text -> plan): bit-exact unit plans, crossfade lengths and pause layout."""
import numpy as np
import pytest


def test_config_loader_matches_reference_defaults_and_shipped(H):
    d = H.front.load_config(None)
    assert (d.crossfade_ms, d.crossfade_vowel_ms, d.word_pause_ms, d.min_silence_ms) == (20.0, 45.0, 120.0, 15.0)
    s = H.shipped_config()
    assert (s.crossfade_ms, s.crossfade_vowel_ms, s.word_pause_ms, s.min_silence_ms) == (90.0, 140.0, 60.0, 35.0)
    assert s.remove_word_silence == 1 and s.remove_dc_offset == 1 and abs(s.default_speed - 1.2) < 1e-6
    if H.have_reference():
        import ctypes as C
        r = H.front.Config()
        H.ref_lib().ctts_load_config.argtypes = [C.POINTER(H.front.Config), C.c_char_p]
        H.ref_lib().ctts_load_config(C.byref(r), H.SHIPPED_YAML.encode())
        assert bytes(r) == bytes(s)


def test_ms_to_samples_known_answers(H, front_small):
    # (size_t)(ms*22050/1000.0f) -> 3ms=66, 60=1323, 35=771
    p = front_small.params()
    assert (p.fade_in_samples, p.min_silence_samples) == (66, 771)
    plan = front_small.plan(["a e, i; o: u. a! e?"])
    ops = plan.ops
    sil = [int(o["a"]) for o in ops if o["kind"] == H.front.OP_SILENCE]
    # word pause 1323; ',' 2381 ';' 2910 ':' 2646 '.' 3969 '!' 4233 '?' 3969
    assert sil == [1323, 2381, 1323, 2910, 1323, 2646, 1323, 3969, 1323, 4233, 1323, 3969]
    assert all(int(o["a"]) == 66 for o in ops if o["kind"] == H.front.OP_FADE_OUT)


def test_rule_count_is_glibc_seven(front_small, golden):
    # glibc rejects the BSD-only [[:<:]] the reference emits: 7 of 49 shipped rules compile
    assert front_small.rule_count == 7 == int(golden["rule_count"][0])


def test_normalized_text_and_unit_traces_match_golden(H, front_small, golden, small_db):
    db = H.voicedb.parse_voice_db(small_db)
    texts = [str(t) for t in golden["e2e_texts"]]
    plan = front_small.plan(texts)
    for k, t in enumerate(texts):
        assert front_small.normalize_text(t) == str(golden[f"e2e_norm_{k}"])
        units = [db.unit_text(int(o["a"])) for o in plan.utt_ops(k) if o["kind"] == H.front.OP_UNIT]
        want = [u for u in golden[f"e2e_units_{k}"].tolist() if u != ""]
        assert units == want, (t, units, want)


def test_number_expansion(front_small):
    f = front_small.normalize_text
    assert f("0") == "zero"
    assert f("21") == "vinte e um"
    assert f("100") == "cem"
    assert f("101") == "cento e um"
    assert f("1100") == "mil cem"
    assert f("1999") == "mil novecentos e noventa e nove"
    assert f("2005") == "dois mil e cinco"
    assert f("1000000") == "um milhão"
    assert f("2000000000") == "dois bilhões"


def test_plan_structure_ola_mundo(H, front_small):
    plan = front_small.plan(["olá mundo"], [1.5])
    ops = plan.utt_ops(0)
    kinds = [int(o["kind"]) for o in ops]
    F = H.front
    # word 1 units, WORD_END, FADE_OUT, SILENCE, MARK, word 2 units, WORD_END, FADE_OUT
    assert kinds[-2:] == [F.OP_WORD_END, F.OP_FADE_OUT]
    assert kinds.count(F.OP_MARK) == 1 and kinds.count(F.OP_WORD_END) == 2
    first = ops[0]
    assert first["kind"] == F.OP_UNIT and first["flags"] & F.UNIT_AFTER_BOUNDARY
    assert float(plan.speed[0]) == 1.5
    assert int(plan.found[0]) == kinds.count(F.OP_UNIT) and int(plan.missing[0]) == 0


def test_unknown_characters_become_silence(H, front_small):
    plan = front_small.plan(["@#"])
    ops = plan.utt_ops(0)
    assert [int(o["kind"]) for o in ops[:2]] == [H.front.OP_SILENCE, H.front.OP_SILENCE]
    assert int(ops[0]["a"]) == 661  # 30 ms
    assert int(plan.missing[0]) == 2


def test_intonation_scalars_by_phrase_type(H, front_small):
    F = H.front
    def word_ends(text):
        return [o for o in front_small.plan([text]).utt_ops(0) if o["kind"] == F.OP_WORD_END]
    # '?' resets word_index, so the circumflex is only reachable when number expansion makes the
    # normalised text longer than the original word count: "21?" is ONE word that becomes three
    q = word_ends("21?")
    assert q[0]["flags"] & F.WE_CIRCUMFLEX and q[0]["flags"] & F.WE_ENERGY
    assert abs(float(q[0]["f2"]) - 1.1) < 1e-6             # peak scaled to the 10 % limit
    assert abs(float(q[0]["e0"]) - 1.05) < 1e-6
    assert not (word_ends("a casa é azul?")[0]["flags"] & F.WE_CIRCUMFLEX)
    d = word_ends("a casa é azul.")
    assert not (d[0]["flags"] & F.WE_ENERGY) and not (d[0]["flags"] & F.WE_CIRCUMFLEX)
    e = word_ends("que casa azul!")
    assert abs(float(e[0]["e0"]) - np.float32(1.25) * np.float32(1.1)) < 1e-6
    assert abs(float(e[0]["f0"]) - 1.1) < 1e-6
    z = word_ends("   ")
    assert all(not (o["flags"] & F.WE_INTON) for o in z)   # total_words == 0


def test_plans_match_reference_unit_traces_live(H, front_small, reference_small, small_db):
    db = H.voicedb.parse_voice_db(small_db)
    texts = H.corpus.batch(40, seed=321)
    plan = front_small.plan(texts)
    for k, t in enumerate(texts):
        units = [db.unit_text(int(o["a"])) for o in plan.utt_ops(k) if o["kind"] == H.front.OP_UNIT]
        assert units == reference_small.unit_trace(t)
        assert front_small.normalize_text(t) == reference_small.normalized_text(t)


def test_bounds_and_select(H, front_small, small_db):
    texts = H.corpus.batch(6, seed=9)
    plan = front_small.plan(texts, [1.0, 1.5, 1.0, 0.5, 2.0, 1.0])
    pre, out, region = front_small.bounds(plan)
    assert (out[[0, 2, 5]] == pre[[0, 2, 5]]).all()
    # WSOLA bound: frames*hop + 512 with hop=(size_t)(128/speed)  (ctts.c:3511-3517)
    frames = (pre[3] - 512) // 128 + 1
    assert out[3] == frames * 256 + 512
    assert (region <= pre).all() and (region > 0).all()
    sub = plan.select([4, 1])
    assert sub.n_utts == 2 and float(sub.speed[0]) == 2.0
    assert np.array_equal(sub.utt_ops(1), plan.utt_ops(1))


def test_threaded_planning_is_identical(H, front_small, monkeypatch):
    """ctts_front_plan_batch plans contiguous slices on worker threads (each with its own compiled
    rules): the concatenated plan must be byte-identical to the single-threaded one."""
    texts = H.corpus.batch(700, seed=11, target_chars=60) + ["", "olá mundo", "12 casas, 3 rios!"]
    speeds = np.linspace(0.5, 2.0, len(texts)).astype(np.float32)
    monkeypatch.setenv("CTTS_FRONT_THREADS", "1")
    a = front_small.plan(texts, speeds)
    for t in ("3", "8"):
        monkeypatch.setenv("CTTS_FRONT_THREADS", t)
        b = front_small.plan(texts, speeds)
        assert np.array_equal(a.utt_op_begin, b.utt_op_begin) and a.ops.tobytes() == b.ops.tobytes()
        assert np.array_equal(a.found, b.found) and np.array_equal(a.missing, b.missing)
        assert np.array_equal(a.speed, b.speed)


def test_zipf_vocabulary_is_seeded_and_skewed(H):
    """corpus.Vocabulary (the vocabulary-size sensitivity of bench.py): deterministic for a seed, the shipped word
    list first, rank-frequency roughly 1 / (rank + 2.7); the default generator is untouched by it."""
    import collections
    c = H.corpus
    v1, v2 = c.Vocabulary(5000, seed=7), c.Vocabulary(5000, seed=7)
    assert v1.words == v2.words and len(set(v1.words)) == 5000
    assert set(v1.words[:len(dict.fromkeys(c.WORDS))]) == set(c.WORDS)
    a = c.batch(300, seed=3, vocab=v1)
    assert a == c.batch(300, seed=3, vocab=v2) and a != c.batch(300, seed=4, vocab=v1)
    counts = collections.Counter(w.strip(".,;:?!").lower() for t in a for w in t.split() if w.strip(".,;:?!-").isalpha())
    top = counts.most_common(1)[0][1] / sum(counts.values())
    assert 0.02 < top < 0.08                      # p(rank 0) = 1 / (2.7 * H) ~ 4 %
    assert sum(1 for n in counts.values() if n == 1) > 0.25 * len(counts)   # a long tail of words seen once
    assert c.batch(5, seed=1234)[0].startswith("Diferente trabalho")        # the benchmark corpus itself did not move
