"""Parity tests proper: the CUDA path through the C-ABI against the oracle (bit-exact), the
committed reference vectors, and size-independent properties at full batch size.

Run on a B200 with:  python -m pytest tests -m gpu
"""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(H):
    return H.importlib.import_module("2026-simple-c-tts_b200.gpu")


@pytest.fixture(scope="module")
def synth_small(gpu, small_db):
    return gpu.GpuSynth(small_db, 0)


def _assert_same(got, want, what=""):
    assert len(got) == len(want), f"{what}: length {len(got)} != {len(want)}"
    if not np.array_equal(got, want):
        d = np.nonzero(got != want)[0]
        raise AssertionError(f"{what}: {len(d)} of {len(want)} samples differ, first at {d[:5]}, "
                             f"max |d| {np.abs(got.astype(int) - want.astype(int)).max()}")


def _check_plan(H, synth, oracle, prm, plan, what=""):
    outs = synth.synth_list(plan, prm)
    for u in range(plan.n_utts):
        want, _ = oracle.synth(prm, plan.utt_ops(u), float(plan.speed[u]))
        _assert_same(outs[u], want, f"{what} utt {u} speed {float(plan.speed[u])}")
    return outs


def test_config1_and_2_ola_mundo(H, synth_small, oracle_small, front_small, golden):
    """BASELINE configs[0] and [1]: 'olá mundo' at 1.0 and at 1.5 (WSOLA), GPU vs oracle (±0)
    and vs the PCM the compiled reference produced (±1 LSB as north_star states; observed ±0)."""
    prm = front_small.params()
    plan = front_small.plan(["olá mundo", "olá mundo"], [1.0, 1.5])
    outs = _check_plan(H, synth_small, oracle_small, prm, plan, "olá mundo")
    for k, got in enumerate(outs):
        want = golden[f"e2e_pcm_{k}"]
        assert len(got) == len(want)
        assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1   # tolerance: ±1 int16 LSB


def test_golden_reference_pcm(H, synth_small, oracle_small, front_small, golden):
    texts = [str(t) for t in golden["e2e_texts"]]
    speeds = [float(s) for s in golden["e2e_speeds"]]
    prm = front_small.params()
    plan = front_small.plan(texts, speeds)
    outs = synth_small.synth_list(plan, prm)
    for k, got in enumerate(outs):
        want = golden[f"e2e_pcm_{k}"]
        assert len(got) == len(want), texts[k]
        _, st = oracle_small.synth(prm, plan.utt_ops(k), 1.0)
        if speeds[k] == 1.0:
            ok = ~H.ub_mask(st, len(got))        # samples the reference computed from out-of-bounds reads
            assert np.abs(got[ok].astype(np.int32) - want[ok].astype(np.int32)).max(initial=0) <= 1
            assert np.array_equal(got[ok], want[ok])
        elif st.ub_spans == 0:
            assert np.array_equal(got, want), texts[k]


def test_batch_speed1_bit_exact(H, synth_small, oracle_small, front_small):
    texts = H.corpus.batch(64, seed=42)
    prm = front_small.params()
    plan = front_small.plan(texts)
    _check_plan(H, synth_small, oracle_small, prm, plan, "batch64")


def test_mixed_speeds_bit_exact(H, synth_small, oracle_small, front_small):
    texts = H.corpus.batch(14, seed=43, target_chars=120)
    speeds = [0.5, 0.6, 0.75, 0.9, 0.995, 1.0, 1.005, 1.1, 1.3, 1.5, 1.7, 2.0, 3.0, 0.2]
    prm = front_small.params()
    plan = front_small.plan(texts, speeds)
    _check_plan(H, synth_small, oracle_small, prm, plan, "mixed")


def test_edge_cases(H, synth_small, oracle_small, front_small):
    texts = ["", " ", "   ", ".", "?!", "a", "@#$", "a-a-a", "...a...", "a,a;a:a.a!a?", "1", "21?", "100 000",
             "á" * 40, "ai " * 30, "(a) [e] \"i\" 'o' `u`", "a\tb\nc\rd", "Dr. Sr. km etc.",
             "casa-casa-casa-casa-casa-casa-casa-casa-casa-casa-casa-casa"]
    prm = front_small.params()
    for sp in (1.0, 1.5):
        plan = front_small.plan(texts, [sp] * len(texts))
        _check_plan(H, synth_small, oracle_small, prm, plan, f"edge@{sp}")


def test_huge_region_uses_hbm_window(H, synth_small, oracle_small, front_small):
    # one "word" far longer than the shared-memory window: commas do not reset the word mark
    text = ",".join(["casa"] * 60)
    prm = front_small.params()
    plan = front_small.plan([text, "olá mundo", text + " fim"], [1.0, 1.0, 1.25])
    pre, _, region = front_small.bounds(plan)
    assert region[0] > 200000
    _check_plan(H, synth_small, oracle_small, prm, plan, "huge")


def test_other_configs(H, gpu, small_db):
    """Compiled-in defaults (crossfade 20 ms, pause 120 ms, ...) and switches off."""
    F = H.front
    texts = H.corpus.batch(10, seed=44, target_chars=100)
    orc = H.Oracle(small_db)
    g = gpu.GpuSynth(small_db, 0)
    for variant in range(4):
        cfg = F.load_config(None)
        if variant == 1:
            cfg.remove_dc_offset = 0
        if variant == 2:
            cfg.remove_word_silence = 0
            cfg.crossfade_ms = 0.0
        if variant == 3:
            cfg.crossfade_ms = 150.0
            cfg.crossfade_vowel_ms = 200.0
            cfg.word_pause_ms = 0.0
            cfg.fade_in_ms = 0.0
            cfg.fade_out_ms = 10.0
            cfg.min_silence_ms = 5.0
            cfg.silence_threshold = 0.2
            cfg.max_pitch_change = 0.3
        fr = F.Front(small_db, cfg, H.NORM_CSV)
        prm = fr.params()
        plan = fr.plan(texts, [1.0] * 9 + [1.4])
        _check_plan(H, g, orc, prm, plan, f"config variant {variant}")


def test_caller_offsets_and_errors(H, gpu, synth_small, oracle_small, front_small):
    prm = front_small.params()
    plan = front_small.plan(H.corpus.batch(3, seed=45, target_chars=60), [1.0, 1.5, 1.0])
    b = synth_small.bounds(plan).astype(np.int64)
    # generous, unevenly padded slots
    off = np.zeros(4, np.uint64)
    off[1] = ((b[0] + 7) // 8) * 8 + 64
    off[2] = off[1] + ((b[1] + 7) // 8) * 8 + 8
    off[3] = off[2] + ((b[2] + 7) // 8) * 8 + 800
    pcm = np.full(int(off[3]), 12345, np.int16)
    _, _, cnt = synth_small.synth_batch(plan, prm, pcm, off)
    for u in range(3):
        want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(plan.speed[u]))
        _assert_same(pcm[int(off[u]):int(off[u]) + int(cnt[u])], want, f"offsets utt {u}")
        assert cnt[u] <= b[u]
    # a slot smaller than its bound is refused, not overrun
    bad = off.copy()
    bad[1] = 8
    with pytest.raises(gpu.GpuError) as e:
        synth_small.synth_batch(plan, prm, pcm, bad)
    assert str(gpu.ERR_BOUNDS) in str(e.value)
    mis = off.copy()
    mis[1] += 3
    with pytest.raises(gpu.GpuError):
        synth_small.synth_batch(plan, prm, pcm, mis)
    # an op naming a unit that does not exist is an invalid plan
    broken = front_small.plan(["a"])
    broken.ops["a"][broken.ops["kind"] == H.front.OP_UNIT] = 10 ** 6
    with pytest.raises(gpu.GpuError):
        synth_small.synth_batch(broken, prm)


def test_resident_plan_is_idempotent(H, synth_small, oracle_small, front_small):
    prm = front_small.params()
    plan = front_small.plan(H.corpus.batch(12, seed=46, target_chars=80), [1.0] * 8 + [0.8, 1.2, 1.6, 2.0])
    rp = synth_small.create_plan(plan, prm)
    rp.run()
    a = rp.utterances()
    rp.run()
    rp.run()
    b = rp.utterances()
    for u in range(plan.n_utts):
        assert np.array_equal(a[u], b[u])
        want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(plan.speed[u]))
        _assert_same(b[u], want, f"resident utt {u}")
    info = rp.info()
    assert info.kernel_launches == 5 and info.n_stretch == 4 and info.threads == 256
    # pre-stretch buffer of a stretched utterance equals the oracle's
    _, _, pre = oracle_small.synth(prm, plan.utt_ops(9), 1.2, want_pre=True)
    _assert_same(rp.read_pre(9, len(pre) + 16), pre, "pre-stretch")


def test_full_voice_sample_against_oracle(H, gpu):
    """The bench voice (1787 units) and corpus: 48 utterances bit-exact, both speeds."""
    db = H.synthetic_db()
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    orc = H.Oracle(db)
    g = gpu.GpuSynth(db, 0)
    texts = H.corpus.batch(48, seed=1234)
    plan = fr.plan(texts, [1.0] * 40 + list(H.corpus.mixed_speeds(8, seed=3)))
    _check_plan(H, g, orc, prm, plan, "full voice")


def test_full_batch_properties(H, gpu):
    """BASELINE configs[2] at full size (4096 utterances): properties that do not need the oracle on
    every utterance -- counts within bounds, idempotence, independence of batch composition (an
    utterance synthesised alone equals the same utterance inside the batch), and 96 utterances bit-exact vs the oracle."""
    db = H.synthetic_db()
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    g = gpu.GpuSynth(db, 0)
    texts = H.corpus.batch(4096, seed=1234)
    plan = fr.plan(texts)
    rp = g.create_plan(plan, prm)
    rp.run()
    cnt = rp.counts().astype(np.int64)
    off = rp.out_offsets().astype(np.int64)
    bounds = g.bounds(plan).astype(np.int64)
    assert (cnt <= bounds).all() and (cnt > 0).all()
    pcm1 = rp.read_pcm(0, rp.out_samples)
    digest1 = [hashlib.sha1(pcm1[off[u]:off[u] + cnt[u]].tobytes()).digest() for u in range(0, 4096, 64)]
    rp.run()
    cnt2 = rp.counts().astype(np.int64)
    assert np.array_equal(cnt, cnt2)
    pcm2 = rp.read_pcm(0, rp.out_samples)
    digest2 = [hashlib.sha1(pcm2[off[u]:off[u] + cnt[u]].tobytes()).digest() for u in range(0, 4096, 64)]
    assert digest1 == digest2
    orc = H.Oracle(db)
    rng = np.random.default_rng(1)
    pick = sorted(rng.choice(4096, 96, replace=False).tolist())
    solo = g.synth_list(plan.select(pick), prm)
    for k, u in enumerate(pick):
        inside = pcm1[off[u]:off[u] + cnt[u]]
        _assert_same(solo[k], inside, f"utt {u} alone vs in batch")
        want, _ = orc.synth(prm, plan.utt_ops(u), 1.0)
        _assert_same(inside, want, f"utt {u} vs oracle")


def test_forced_hbm_window_and_chunked_batch(H, gpu, small_db, oracle_small, front_small, monkeypatch):
    """The two rarely taken execution paths, forced: (a) a shared window so small that most regions
    are assembled in the HBM slot and joins reach back across tasks (CTTS_GPU_WINDOW), (b) the drop-in
    call cut into many launches whose copies overlap the next launch (CTTS_GPU_CHUNK_SAMPLES)."""
    prm = front_small.params()
    texts = H.corpus.batch(24, seed=47, target_chars=120) + ["", "a", ". . .", "olá mundo"]
    plan = front_small.plan(texts)
    want = [oracle_small.synth(prm, plan.utt_ops(u), 1.0)[0] for u in range(plan.n_utts)]
    # (c) a unit-head pitch table with room for 3 entries: every other join estimates both signals itself
    for env in ({"CTTS_GPU_WINDOW": "2048"}, {"CTTS_GPU_CHUNK_SAMPLES": "300000"},
                {"CTTS_GPU_WINDOW": "4096", "CTTS_GPU_CHUNK_SAMPLES": "100000", "CTTS_GPU_CTAS_PER_SM": "2"},
                {"CTTS_GPU_PITCH_SLOTS": "3"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        g = gpu.GpuSynth(small_db, 0)
        outs = g.synth_list(plan, prm)
        for u in range(plan.n_utts):
            _assert_same(outs[u], want[u], f"{env} utt {u}")
        rp = g.create_plan(plan, prm)
        rp.run()
        for u, got in enumerate(rp.utterances()):
            _assert_same(got, want[u], f"{env} resident utt {u}")
        if "CTTS_GPU_WINDOW" in env:
            assert rp.info().n_global_tasks > 0
        for k in env:
            monkeypatch.delenv(k)


def test_bounds_are_upper_bounds_and_tighter_than_the_sum(H, synth_small, oracle_small, front_small):
    """ctts_gpu_plan_bounds subtracts the crossfade overlap every certain join must consume; it must
    still bound the true counts (pre-stretch and output) for every utterance, at every speed."""
    prm = front_small.params()
    texts = H.corpus.batch(40, seed=48, target_chars=90)
    speeds = [1.0] * 30 + [0.5, 0.7, 0.9, 1.1, 1.3, 1.5, 1.7, 1.9, 2.0, 0.6]
    plan = front_small.plan(texts, speeds)
    tight = synth_small.bounds(plan).astype(np.int64)
    _, loose, _ = front_small.bounds(plan)
    assert (tight <= loose.astype(np.int64)).all() and (tight[:30] < loose[:30].astype(np.int64)).any()
    for u in range(plan.n_utts):
        pcm, st = oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))
        assert len(pcm) <= tight[u], (u, len(pcm), tight[u])


def test_two_resident_plans_with_different_windows(H, gpu, synth_small, oracle_small, front_small):
    """Plans whose shared-memory windows differ (a tiny batch next to a large region) stay runnable in
    any order: the dynamic shared-memory limit is a per-device setting, not a per-plan one."""
    prm = front_small.params()
    small = front_small.plan(["a", "olá"])
    big = front_small.plan([",".join(["casa"] * 12), "bom dia mundo"])
    ra = synth_small.create_plan(big, prm)
    rb = synth_small.create_plan(small, prm)
    assert ra.info().smem_bytes != rb.info().smem_bytes
    for rp, plan in ((ra, big), (rb, small), (ra, big)):
        rp.run()
        for u, got in enumerate(rp.utterances()):
            want, _ = oracle_small.synth(prm, plan.utt_ops(u), 1.0)
            _assert_same(got, want, f"utt {u}")


def test_mixed_speed_batch_properties(H, gpu):
    """BASELINE configs[3] at FULL size (4096 utterances of ~200 characters, speeds 0.5-2.0): counts within
    bounds, idempotence, known-answer WSOLA lengths (frames, hop and the trailing-zero trim of
    ctts.c:3515-3517, :3611), and a bit-exact oracle check of 64 utterances spread over the batch."""
    db = H.synthetic_db()
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    g = gpu.GpuSynth(db, 0)
    n = 4096
    texts = H.corpus.batch(n, seed=4321, target_chars=200)
    speeds = H.corpus.mixed_speeds(n, seed=5)
    plan = fr.plan(texts, speeds)
    rp = g.create_plan(plan, prm)
    rp.run()
    cnt = rp.counts().astype(np.int64)
    off = rp.out_offsets().astype(np.int64)
    assert (cnt <= g.bounds(plan).astype(np.int64)).all() and (cnt > 0).all()
    digest1 = hashlib.sha1(rp.read_pcm(0, rp.out_samples).tobytes()).hexdigest()
    rp.run()
    assert np.array_equal(cnt, rp.counts().astype(np.int64))
    assert digest1 == hashlib.sha1(rp.read_pcm(0, rp.out_samples).tobytes()).hexdigest()
    st = rp.wsola_stats()
    # the speculated offsets hold on this workload: no chain is walked, tier 1 rejects almost every
    # candidate (65 + 6 per frame) and the exact loop is rare
    assert st.frames > 0 and st.walked_utterances == 0 and st.walked_frames == 0
    assert st.tier2_candidates < 2 * st.frames and st.exact_evaluations < 0.05 * st.frames
    orc = H.Oracle(db)
    for u in range(7, n, 64):      # 64 utterances
        want, ost, pre = orc.synth(prm, plan.utt_ops(u), float(speeds[u]), want_pre=True)
        # length known-answer: (frames-1)*hop + 512 minus trailing zeros
        n_frames = (len(pre) - 512) // 128 + 1
        hop = int(128 / np.float32(speeds[u]))
        assert ost.wsola_frames == n_frames and len(want) <= (n_frames - 1) * hop + 512
        _assert_same(rp.read_pcm(int(off[u]), int(cnt[u])), want, f"utt {u} speed {float(speeds[u])}")


def test_paragraph_batch_properties(H, gpu):
    """BASELINE configs[4] (paragraph-length utterances, ~30 s of audio each, prosody / pitch smoothing over
    many words and joins), one GPU's worth of it at test size: 1024 paragraphs through the drop-in call in
    several pieces; counts within bounds, every phrase type present, 64 utterances bit-exact vs the oracle."""
    db = H.synthetic_db()
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    g = gpu.GpuSynth(db, 0)
    n = 1024
    texts = H.corpus.batch(n, seed=8642, target_chars=215)
    plan = fr.plan(texts)
    pcm, off, cnt = g.synth_batch(plan, prm)
    assert (cnt.astype(np.int64) <= g.bounds(plan).astype(np.int64)).all()
    secs = cnt.astype(np.float64) / 22050
    assert secs.mean() > 25 and {t.rstrip()[-1] for t in texts} >= {".", "?", "!"}
    orc = H.Oracle(db)
    for u in range(3, n, 16):      # 64 utterances
        want, _ = orc.synth(prm, plan.utt_ops(u), 1.0)
        _assert_same(pcm[int(off[u]):int(off[u]) + int(cnt[u])], want, f"paragraph {u}")


def test_reference_hashes_of_corpus_sentences(H, gpu, synth_small, front_small, oracle_small, small_db):
    """The CUDA path against the COMPILED REFERENCE on the GPU box (where the reference tree is absent): SHA-1 +
    length of the reference's PCM for 48 corpus sentences (tests/golden/ref_corpus.json, make_golden_corpus.py),
    through the text -> PCM entry point.  Samples tainted by the reference's out-of-bounds read are zeroed on both sides."""
    pipe = H.importlib.import_module("2026-simple-c-tts_b200.pipeline")
    rows = H.golden_corpus()["corpus"]
    prm = front_small.params()
    texts, speeds = [r["text"] for r in rows], [r["speed"] for r in rows]
    plan = front_small.plan(texts, speeds)
    batch = pipe.TextBatch(texts, speeds)
    pcm = np.zeros(pipe.capacity_hint(front_small, batch), dtype=np.int16)
    off, cnt, _, _ = pipe.synth_texts(front_small, synth_small, batch, pcm)
    for u, r in enumerate(rows):
        _, st = oracle_small.synth(prm, plan.utt_ops(u), r["speed"])      # only for the taint mask
        got = pcm[int(off[u]):int(off[u]) + int(cnt[u])]
        assert len(got) == r["samples"], r["text"]
        assert H.masked_sha1(got, H.ub_mask(st, len(got))) == r["sha1"], (u, r["text"], r["speed"])


def test_voice_with_an_odd_pcm_offset(H, gpu):
    """ctts_gpu_init on a voice.db whose PCM pool starts at an odd byte offset (ctts.c:1001-1004, :1159; the
    loader re-packs it with byte copies) + the front end on the same bytes: the compiled reference's hashes."""
    db = H.odd_offset_voice()
    assert H.voicedb.parse_voice_db(db).audio_offset % 2 == 1
    rows = H.golden_corpus()["odd_voice"]
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    orc = H.Oracle(db)
    g = gpu.GpuSynth(db, 0)
    plan = fr.plan([r["text"] for r in rows], [r["speed"] for r in rows])
    outs = g.synth_list(plan, prm)
    for u, r in enumerate(rows):
        want, st = orc.synth(prm, plan.utt_ops(u), r["speed"])
        _assert_same(outs[u], want, r["text"])
        assert len(outs[u]) == r["samples"] and H.masked_sha1(outs[u], H.ub_mask(st, len(outs[u]))) == r["sha1"], r["text"]


def test_bounded_random_stress(H, gpu):
    """A bounded version of tools/stress_parity.py inside the suite: 2 configurations x 150 random utterances
    (lengths 5-260 characters, 30 of them at random speeds 0.45-2.1), every one bit-exact vs the oracle."""
    db = H.synthetic_db()
    rng = np.random.default_rng(20261018)
    orc = H.Oracle(db)
    g = gpu.GpuSynth(db, 0)
    c2 = H.shipped_config()
    c2.remove_dc_offset = 0
    c2.min_silence_ms = 12.0
    c2.silence_threshold = 0.1
    c2.crossfade_ms = 150.0
    c2.word_pause_ms = 5.0
    for cfg in (H.shipped_config(), c2):
        fr = H.front.Front(db, cfg, H.NORM_CSV)
        prm = fr.params()
        lens = rng.integers(5, 260, size=150)
        texts = [H.corpus.sentence(rng, int(L)) for L in lens]
        speeds = np.ones(150, dtype=np.float32)
        speeds[120:] = rng.uniform(0.45, 2.1, size=30).astype(np.float32)
        plan = fr.plan(texts, speeds)
        outs = g.synth_list(plan, prm)
        for u in range(plan.n_utts):
            want, _ = orc.synth(prm, plan.utt_ops(u), float(speeds[u]))
            _assert_same(outs[u], want, f"stress utt {u} speed {float(speeds[u])} {texts[u][:40]!r}")


def test_chunked_stretch_batch(H, gpu, small_db, oracle_small, front_small, monkeypatch):
    """ctts_gpu_synth_batch on a batch that mixes stretched and plain utterances, cut into several
    launches of assemble -> WSOLA (scan, verify, chain walk, overlap-add) whose device->host copies overlap
    the next chunk (CTTS_GPU_CHUNK_SAMPLES forces small chunks); the result must not depend on the cut."""
    prm = front_small.params()
    texts = H.corpus.batch(22, seed=53, target_chars=100) + ["", "olá mundo", "a"]
    speeds = [1.5, 1.0, 0.5, 2.0, 1.0, 0.7, 1.3, 1.0, 0.9, 1.1, 1.0, 1.7, 0.6, 1.0, 1.9, 0.8,
              1.0, 1.2, 1.4, 1.0, 1.6, 0.5, 1.5, 1.5, 1.0]
    plan = front_small.plan(texts, speeds)
    want = [oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))[0] for u in range(plan.n_utts)]
    for chunk in ("1", "300000", "1000000", "100000000"):
        monkeypatch.setenv("CTTS_GPU_CHUNK_SAMPLES", chunk)
        g = gpu.GpuSynth(small_db, 0)
        outs = g.synth_list(plan, prm)
        for u in range(plan.n_utts):
            _assert_same(outs[u], want[u], f"chunk {chunk} utt {u} speed {speeds[u]}")
    monkeypatch.delenv("CTTS_GPU_CHUNK_SAMPLES")


def test_wsola_chain_walk_from_unverified_frames(H, gpu, small_db, oracle_small, front_small, monkeypatch):
    """The WSOLA frame chain (ctts.c:3555-3592) is speculated and verified frame by frame; where the
    verification fails the chain is walked from that frame on (wsola_search_kernel).  Forced here:
    every N-th frame reported as unverified (the walk starts in the middle of an utterance, with the
    verified position before it), and speculation switched off (every chain walked from frame 1).
    The PCM must not change."""
    prm = front_small.params()
    texts = H.corpus.batch(10, seed=97, target_chars=90) + ["olá mundo", "a", ""]
    speeds = [1.5, 0.5, 2.0, 0.7, 1.3, 0.9, 1.1, 1.7, 0.6, 1.9, 1.5, 1.5, 1.5]
    plan = front_small.plan(texts, speeds)
    want = [oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))[0] for u in range(plan.n_utts)]
    for env, val in (("CTTS_GPU_WSOLA_FORCE_BAD", "1"), ("CTTS_GPU_WSOLA_FORCE_BAD", "2"), ("CTTS_GPU_WSOLA_FORCE_BAD", "37"),
                     ("CTTS_GPU_WSOLA_FORCE_BAD", "300"), ("CTTS_GPU_WSOLA_SPECULATE", "0")):
        monkeypatch.setenv(env, val)
        g = gpu.GpuSynth(small_db, 0)
        rp = g.create_plan(plan, prm)
        rp.run()
        outs = rp.utterances()
        for u in range(plan.n_utts):
            _assert_same(outs[u], want[u], f"{env}={val} utt {u} speed {speeds[u]}")
        st = rp.wsola_stats()
        assert st.walked_utterances >= 10 and st.walked_frames > 0, (env, val, st.walked_utterances)
        if env == "CTTS_GPU_WSOLA_SPECULATE":
            assert st.walked_frames >= st.frames - plan.n_utts and st.tier2_candidates == 0
        monkeypatch.delenv(env)
    g = gpu.GpuSynth(small_db, 0)
    rp = g.create_plan(plan, prm)
    rp.run()
    assert rp.wsola_stats().walked_utterances == 0


def test_wsola_ties_on_an_exactly_periodic_voice(H, gpu):
    """A voice whose units are exactly periodic (period 64 samples, no noise): inside a unit the candidate
    windows one and two periods before the speculated one are the same samples, they score exactly 1.0f
    too and win by scan order (ctts.c:3459 keeps the first maximum), so the speculated offsets are wrong
    and the verification must notice: those utterances are resolved by the chain walk.  Bit-exact vs the oracle."""
    rng = np.random.default_rng(5)
    units = []
    for ch in "abcdelmnoprstuv":
        pattern = rng.integers(-9000, 9000, size=64)
        n = int(rng.integers(3000, 6000))
        units.append((ch, np.tile(pattern, n // 64 + 1)[:n].astype(np.int16)))
    db = H.voicedb.build_voice_db(units)
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    texts = ["a", "aba", "casa de pedra", "um pote de mel", "la no alto da serra, a lua nova"]
    speeds = [1.5, 0.8, 1.7, 0.6, 1.3]
    plan = fr.plan(texts, speeds)
    orc = H.Oracle(db)
    g = gpu.GpuSynth(db, 0)
    rp = g.create_plan(plan, prm)
    rp.run()
    outs = rp.utterances()
    for u in range(plan.n_utts):
        want, _ = orc.synth(prm, plan.utt_ops(u), float(speeds[u]))
        _assert_same(outs[u], want, f"periodic voice utt {u} speed {speeds[u]}")
    st = rp.wsola_stats()
    assert st.walked_utterances >= 2 and st.exact_evaluations > 0


def test_target_rms_variants_share_a_context(H, synth_small, oracle_small, front_small):
    """normalize_rms (ctts.c:1709) lives in a per-context table keyed by target_rms (the normalized
    pool and the unit-head pitch table): plans with different targets -- off, gain clamped at 3.0
    and at 0.1 -- alternate on one context and each must match the oracle bit for bit."""
    import copy
    texts = H.corpus.batch(10, seed=61, target_chars=110) + ["olá mundo"]
    speeds = [1.0] * 9 + [1.5, 1.0]
    plan = front_small.plan(texts, speeds)
    base = front_small.params()
    plans = []
    for target in (3000.0, 0.0, 20000.0, 150.0, 3000.0):
        prm = copy.copy(base)
        prm.target_rms = target
        plans.append((prm, synth_small.create_plan(plan, prm)))
    for prm, rp in plans + plans[::-1]:
        rp.run()
        for u, got in enumerate(rp.utterances()):
            want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))
            _assert_same(got, want, f"target_rms {prm.target_rms} utt {u}")


def test_many_runs_of_one_plan_are_identical(H, gpu):
    """A race between CTAs or between the phases of one CTA shows up as a run that differs from the
    others (or as a device fault), rarely: 40 runs of a 512-utterance resident plan must all produce
    the same counts and the same PCM (compared through a 64-bit sum and a strided sample)."""
    db = H.synthetic_db()
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    g = gpu.GpuSynth(db, 0)
    plan = fr.plan(H.corpus.batch(512, seed=777, target_chars=200))
    rp = g.create_plan(plan, prm)
    first = None
    for k in range(40):
        rp.run()
        cnt = rp.counts().astype(np.int64)
        pcm = rp.read_pcm(0, rp.out_samples)
        sig = (int(cnt.sum()), int(pcm.astype(np.int64).sum()), pcm[::4099].tobytes())
        if first is None:
            first = sig
        assert sig == first, f"run {k} differs from run 0"


def test_many_calls_of_the_drop_in_path_are_identical(H, gpu, monkeypatch):
    """The same for ctts_gpu_synth_batch on a batch that mixes speeds: chunks on several streams,
    per-chunk copies, arenas reused from call to call.  12 calls must return identical PCM."""
    monkeypatch.setenv("CTTS_GPU_CHUNK_SAMPLES", "4000000")   # several chunks in flight
    db = H.synthetic_db()
    fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
    prm = fr.params()
    g = gpu.GpuSynth(db, 0)
    texts = H.corpus.batch(160, seed=909, target_chars=160)
    speeds = H.corpus.mixed_speeds(160, seed=11)
    speeds[::5] = 1.0
    plan = fr.plan(texts, speeds)
    first = None
    for k in range(12):
        outs = g.synth_list(plan, prm)
        sig = [(len(o), int(o.astype(np.int64).sum()), o[::997].tobytes()) for o in outs]
        if first is None:
            first = sig
        assert sig == first, f"call {k} differs from call 0"


def test_c_command_line_writes_the_reference_wav(H, small_db, golden, tmp_path):
    """ctts_b200 (plain C over the two C-ABI libraries, the binding a CTTS maintainer would link):
    `synth` at 1.0 and 1.5 must write byte for byte the WAV the compiled reference writes (44-byte header
    of ctts_write_wav, ctts.c:809, + its PCM: BASELINE configs[0] and [1]); `synth-batch` the same files."""
    import shutil
    import struct
    import subprocess
    b = H.importlib.import_module("2026-simple-c-tts_b200._build")
    exe = b.build_cli()
    shutil.copy(H.SHIPPED_YAML, tmp_path / "config.yaml")
    shutil.copy(H.NORM_CSV, tmp_path / "normalization.csv")
    (tmp_path / "voice.db").write_bytes(small_db)
    for k, speed in enumerate(("1.0", "1.5")):
        r = subprocess.run([exe, "synth", "voice.db", "olá mundo", f"o{k}.wav", speed], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        raw = (tmp_path / f"o{k}.wav").read_bytes()
        want = golden[f"e2e_pcm_{k}"]
        assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data" and len(raw) == 44 + 2 * len(want)
        assert struct.unpack("<IHHIIHH", raw[16:36]) == (16, 1, 1, 22050, 44100, 2, 16)
        assert struct.unpack("<I", raw[4:8])[0] == 36 + 2 * len(want) and struct.unpack("<I", raw[40:44])[0] == 2 * len(want)
        assert np.array_equal(np.frombuffer(raw[44:], dtype="<i2"), want)
        assert f"Synthesized {len(want)} samples" in r.stdout and "missing: 0" in r.stdout
    (tmp_path / "t.tsv").write_text("1.0\tolá mundo\n1.5\tolá mundo\n", encoding="utf-8")
    r = subprocess.run([exe, "synth-batch", "voice.db", "t.tsv", "out"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (tmp_path / "out" / "000000.wav").read_bytes() == (tmp_path / "o0.wav").read_bytes()
    assert (tmp_path / "out" / "000001.wav").read_bytes() == (tmp_path / "o1.wav").read_bytes()


def test_streaming_call_hands_over_finished_ranges(H, gpu, small_db, oracle_small, front_small, monkeypatch):
    """ctts_gpu_synth_batch_stream: the callback sees every utterance exactly once, in order, and what it
    finds in the host buffer at that moment already is the final PCM (checked against the oracle);
    plain and stretched utterances, several chunks."""
    monkeypatch.setenv("CTTS_GPU_CHUNK_SAMPLES", "150000")
    prm = front_small.params()
    texts = H.corpus.batch(20, seed=71, target_chars=110) + ["", "olá mundo"]
    for speeds in ([1.0] * 22, [1.5, 1.0, 0.5, 2.0, 1.0, 0.7] * 3 + [1.3, 1.0, 0.9, 1.0]):
        plan = front_small.plan(texts, speeds)
        g = gpu.GpuSynth(small_db, 0)
        seen, snap = [], {}

        def on_chunk(pcm, off, cnt, b, e):
            seen.append((b, e))
            for u in range(b, e):
                snap[u] = pcm[int(off[u]):int(off[u]) + int(cnt[u])].copy()

        pcm, off, cnt = g.synth_batch_stream(plan, prm, on_chunk)
        assert len(seen) > 1 and seen[0][0] == 0 and seen[-1][1] == plan.n_utts
        assert all(seen[i][1] == seen[i + 1][0] for i in range(len(seen) - 1))
        for u in range(plan.n_utts):
            want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))
            _assert_same(snap[u], want, f"utt {u} at callback time")
            _assert_same(pcm[int(off[u]):int(off[u]) + int(cnt[u])], want, f"utt {u} at return")


def test_session_pieces_equal_one_batch(H, gpu, synth_small, oracle_small, front_small):
    """ctts_gpu_session_*: a batch fed as pieces of different sizes (more pieces than lanes, an empty piece,
    plain and stretched utterances) lands in packed slots and equals the oracle bit for bit."""
    prm = front_small.params()
    texts = H.corpus.batch(23, seed=131, target_chars=90) + ["", "olá mundo"]
    speeds = [1.0, 1.5, 1.0, 0.6, 1.0, 2.0, 1.0, 1.0, 0.9, 1.0] * 2 + [1.3, 1.0, 1.0, 1.0, 1.5]
    plan = front_small.plan(texts, speeds)
    cuts = [0, 1, 4, 4, 9, 10, 17, 24, 25]   # (4, 4): an empty piece
    pieces = [plan.select(range(a, b)) for a, b in zip(cuts[:-1], cuts[1:])]
    cap = int(synth_small.layout(plan)[-1])
    pcm = np.zeros(cap + 64, dtype=np.int16)
    off, cnt = synth_small.synth_pieces(pieces, prm, pcm)
    # packed: every utterance right behind the one before it, rounded up to 8 samples
    assert len(off) == plan.n_utts and off[0] == 0
    assert np.array_equal(np.diff(off.astype(np.int64)), (cnt[:-1].astype(np.int64) + 7) // 8 * 8)
    for u in range(plan.n_utts):
        want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))
        _assert_same(pcm[int(off[u]):int(off[u]) + int(cnt[u])], want, f"session utt {u}")
    # a buffer that is too small is an argument error, not a crash
    with pytest.raises(gpu.GpuError):
        synth_small.synth_pieces(pieces, prm, np.zeros(cap // 2, dtype=np.int16))
    # ... and the context is usable afterwards
    off2, cnt2 = synth_small.synth_pieces(pieces[:3], prm, pcm)
    assert np.array_equal(cnt2, cnt[:4])


def test_text_pipeline_equals_the_planned_batch(H, gpu, synth_small, front_small, oracle_small):
    """ctts_b200_synth_texts (planner threads feeding a device session, the text -> PCM drop-in for N calls of
    ctts_synthesize, ctts.c:3623): for every piece size / thread count the PCM, the counts and the unit
    statistics equal those of the batch planned in one go and run through ctts_gpu_synth_batch -- and the oracle."""
    pipe = H.importlib.import_module("2026-simple-c-tts_b200.pipeline")
    prm = front_small.params()
    texts = H.corpus.batch(61, seed=977, target_chars=100) + ["", "olá mundo", "São 1234 casas, não é?"]
    speeds = list(H.corpus.mixed_speeds(64, seed=3))
    speeds[::3] = [1.0] * len(speeds[::3])
    plan = front_small.plan(texts, speeds)
    ref = synth_small.synth_list(plan, prm)
    for u in (0, 7, 33, 62, 63):
        want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))
        _assert_same(ref[u], want, f"utt {u}")
    batch = pipe.TextBatch(texts, speeds)
    pcm = np.zeros(pipe.capacity_hint(front_small, batch), dtype=np.int16)
    cache = pipe.PlanCache(64 << 20)
    for k, (piece_utts, threads) in enumerate(((1, 3), (7, 1), (16, 4), (0, 0), (1000, 2), (0, 0), (5, 3))):
        # the last two runs go through a plan cache: cold (every text planned and stored), then warm (no text planned)
        off, cnt, used, tm, stats = pipe.synth_texts(front_small, synth_small, batch, pcm, piece_utts, threads, want_stats=True,
                                                     cache=cache if k >= 5 else None)
        if k == 5:
            assert cache.stats()["hits"] <= 2 and cache.stats()["entries"] >= plan.n_utts - 2      # ("olá mundo" may repeat)
        if k == 6:
            st = cache.stats()
            assert st["hits"] >= plan.n_utts and st["misses"] <= plan.n_utts
        assert used <= pcm.size and tm.done_s >= tm.all_submitted_s >= 0
        assert np.array_equal(stats[:, 0], plan.found) and np.array_equal(stats[:, 1], plan.missing)
        for u in range(plan.n_utts):
            _assert_same(pcm[int(off[u]):int(off[u]) + int(cnt[u])], ref[u], f"pieces of {piece_utts}, {threads} threads, utt {u}")


def _multi_check(H, gpu, devices, small_db, oracle_small, front_small):
    prm = front_small.params()
    texts = H.corpus.batch(37, seed=211, target_chars=90) + ["", "olá mundo", "a"]
    speeds = list(H.corpus.mixed_speeds(40, seed=17))
    speeds[::2] = [1.0] * len(speeds[::2])
    plan = front_small.plan(texts, speeds)
    ctxs = [gpu.GpuSynth(small_db, d) for d in devices]
    pcm, off, cnt, shard_of = gpu.multi_synth_batch(ctxs, plan, prm)
    assert set(shard_of.tolist()) == set(range(len(devices)))          # every device got work
    slot = np.diff(off.astype(np.int64))
    loads = [int(slot[shard_of == d].sum()) for d in range(len(devices))]
    assert max(loads) - min(loads) <= 3 * int(slot.max())             # LPT balance (costs weigh WSOLA up)
    for u in range(plan.n_utts):
        want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))
        _assert_same(pcm[int(off[u]):int(off[u]) + int(cnt[u])], want, f"multi utt {u} on device {shard_of[u]}")
    # the same contexts are usable one by one afterwards
    one = ctxs[-1].synth_list(plan, prm)
    assert all(np.array_equal(one[u], pcm[int(off[u]):int(off[u]) + int(cnt[u])]) for u in range(plan.n_utts))


def test_multi_device_entry_two_contexts_on_one_gpu(H, gpu, small_db, oracle_small, front_small):
    """ctts_gpu_multi_synth_batch (BASELINE's multi-GPU configuration: one batch sharded by utterance, host-side
    gather into the caller's buffer): partition, per-shard host threads and scattered slots, exercised with
    three contexts on device 0 so that it runs on a one-GPU box."""
    _multi_check(H, gpu, [0, 0, 0], small_db, oracle_small, front_small)


def test_multi_device_entry_on_two_gpus(H, gpu, small_db, oracle_small, front_small):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    _multi_check(H, gpu, list(range(min(torch.cuda.device_count(), 8))), small_db, oracle_small, front_small)


def test_packed_batch_call(H, gpu, synth_small, oracle_small, front_small, monkeypatch):
    """ctts_gpu_synth_batch_packed: the batch call with a library-chosen, packed layout (device prefix sum of the
    counts + gather; exactly the samples that exist cross PCIe), several pieces, plain and stretched utterances."""
    monkeypatch.setenv("CTTS_GPU_CHUNK_SAMPLES", "2000000")
    g = gpu.GpuSynth(small_db_bytes(H), 0)
    prm = front_small.params()
    texts = H.corpus.batch(30, seed=333, target_chars=90) + ["", "olá mundo"]
    speeds = [1.0, 1.5, 1.0, 0.6, 1.0, 2.0, 1.0, 1.0] * 4
    plan = front_small.plan(texts, speeds)
    cap = int(g.layout(plan)[-1])
    pcm = np.full(cap, 12345, dtype=np.int16)
    off, cnt, used = g.synth_batch_packed(plan, prm, pcm)
    assert off[0] == 0 and np.array_equal(np.diff(off.astype(np.int64)), (cnt[:-1].astype(np.int64) + 7) // 8 * 8)
    assert used == int(off[-1]) + (int(cnt[-1]) + 7) // 8 * 8 and used < cap
    for u in range(plan.n_utts):
        want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))
        _assert_same(pcm[int(off[u]):int(off[u]) + int(cnt[u])], want, f"packed utt {u}")
        assert not pcm[int(off[u]) + int(cnt[u]):int(off[u]) + (int(cnt[u]) + 7) // 8 * 8].any()      # zero padding
    assert (pcm[used:] == 12345).all()                                                        # nothing written past the end
    with pytest.raises(gpu.GpuError):
        g.synth_batch_packed(plan, prm, np.zeros(used - 8, dtype=np.int16))


def test_whole_task_copies_when_a_fade_out_reaches_back(H, gpu, small_db, oracle_small):
    """Second level of the deduplication with fade-outs longer than short words (fade_out_ms = 150: 3307 samples): a
    source task whose fade-out reaches back into the previous word does not share its samples (its result depends on
    more than its own ops) and the tasks that were to copy it fall back to their canonical region / assemble
    themselves.  Long crossfades and no word pause move the thresholds as well.  Bit-exact against the oracle."""
    for variant in range(2):
        cfg = H.shipped_config()
        cfg.fade_out_ms = 150.0
        if variant == 1:
            cfg.word_pause_ms = 0.0
            cfg.crossfade_ms = 120.0
            cfg.crossfade_vowel_ms = 160.0
        fr = H.front.Front(small_db, cfg, H.NORM_CSV)
        prm = fr.params()
        base = ["a e o a e o casa a e o", "o a o a o a mundo o a", "e casa e casa e casa e casa", "a o e bom dia a o e bom dia"]
        texts = base * 3 + ["casa " + base[0], base[1] + " dia"]
        speeds = [1.0] * len(texts)
        speeds[5] = 0.6
        plan = fr.plan(texts, speeds)
        g = gpu.GpuSynth(small_db, 0)
        rp = g.create_plan(plan, prm)
        info = rp.info()
        assert info.n_source_tasks >= 4 and info.n_reuse_tasks >= 8, (info.n_source_tasks, info.n_reuse_tasks)
        for _ in range(2):
            rp.run()
        outs = rp.utterances()
        for u in range(plan.n_utts):
            want, _ = oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))
            _assert_same(outs[u], want, f"variant {variant} utt {u} {texts[u]!r}")
        rp.close()
        g.close()


def small_db_bytes(H):
    return H.small_db()


def test_region_dedup_on_and_off(H, gpu, small_db, oracle_small, front_small, monkeypatch):
    """Word-region deduplication: equal word regions of a batch are assembled once per launch and copied by their
    other occurrences, which resume at their own contour.  A batch full of repeated words (and of things that must
    NOT be deduplicated: utterance starts, punctuation-only pieces, hyphens, numbers, stretched utterances) gives
    the same PCM with the deduplication on and off, and the oracle's."""
    prm = front_small.params()
    words = ["casa", "mundo", "olá", "bom", "dia", "rato", "rua", "importante", "água", "trabalho"]
    rng = np.random.default_rng(12)
    texts = [" ".join(rng.choice(words, size=int(rng.integers(3, 14)))) + rng.choice([".", "?", "!", ",", ""]) for _ in range(40)]
    texts += ["casa, casa; casa: casa. casa! casa?", "a casa casa-casa casa", "12 casas e 12 casas", "...", "", "casa", "casa casa"]
    texts += [texts[0], texts[0], texts[5], texts[11], "bom dia mundo", "bom dia mundo", "olá bom dia mundo"]   # equal whole tasks
    speeds = [1.0] * len(texts)
    speeds[3] = 1.5
    speeds[11] = 0.7
    speeds[-4] = 0.7
    plan = front_small.plan(texts, speeds)
    want = [oracle_small.synth(prm, plan.utt_ops(u), float(speeds[u]))[0] for u in range(plan.n_utts)]
    outs = {}
    for mode in ("2", "1", "0"):
        monkeypatch.setenv("CTTS_GPU_REGION_DEDUP", mode)
        g = gpu.GpuSynth(small_db, 0)
        rp = g.create_plan(plan, prm)
        info = rp.info()
        if mode != "0":
            assert info.n_canon_tasks >= 5 and info.n_dedup_tasks > 4 * info.n_canon_tasks, (info.n_canon_tasks, info.n_dedup_tasks)
        else:
            assert info.n_canon_tasks == 0 and info.n_dedup_tasks == 0
        if mode == "2":      # second level: whole tasks that are equal (same region, same contour, same pause)
            assert info.n_source_tasks >= 10 and info.n_reuse_tasks >= info.n_source_tasks, (info.n_source_tasks, info.n_reuse_tasks)
        else:
            assert info.n_source_tasks == 0 and info.n_reuse_tasks == 0
        for _ in range(2):          # a second run of the same plan: a new epoch of the region states
            rp.run()
        outs[mode] = rp.utterances()
        for u in range(plan.n_utts):
            _assert_same(outs[mode][u], want[u], f"dedup {mode} utt {u} {texts[u][:30]!r}")
        # the drop-in call (pieces) too
        piece_outs = g.synth_list(plan, prm)
        for u in range(plan.n_utts):
            _assert_same(piece_outs[u], want[u], f"dedup {mode}, pieces, utt {u}")
    monkeypatch.delenv("CTTS_GPU_REGION_DEDUP")
