"""Randomised large-sample parity check (GPU vs oracle, bit-exact) beyond what the test-suite runs:
python tools/stress_parity.py [n_speed1] [n_stretched] [seed]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
gpu = importlib.import_module("2026-simple-c-tts_b200.gpu")
n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 200
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 777
rng = np.random.default_rng(seed)
db = H.synthetic_db()
bad = 0
t0 = time.time()
def variants():
    yield H.shipped_config()
    yield H.front.load_config(None)
    c = H.shipped_config(); c.remove_dc_offset = 0; c.min_silence_ms = 12.0; c.silence_threshold = 0.1; yield c
    c = H.front.load_config(None); c.remove_word_silence = 0; c.crossfade_ms = 0.0; c.max_pitch_change = 0.3; yield c
    c = H.shipped_config(); c.crossfade_ms = 150.0; c.crossfade_vowel_ms = 200.0; c.word_pause_ms = 5.0; c.fade_out_ms = 20.0; yield c
    c = H.shipped_config(); c.word_pause_ms = 0.0; c.unknown_silence_ms = 0.0; c.fade_in_ms = 0.0; c.fade_out_ms = 0.0; yield c


for variant, cfg in enumerate(variants()):
    fr = H.front.Front(db, cfg, H.NORM_CSV)
    prm = fr.params()
    orc = H.Oracle(db)
    g = gpu.GpuSynth(db, 0)
    lens = rng.integers(5, 260, size=n1 + n2)
    texts = [H.corpus.sentence(rng, int(L)) for L in lens]
    speeds = np.ones(n1 + n2, dtype=np.float32)
    speeds[n1:] = rng.uniform(0.45, 2.1, size=n2).astype(np.float32)
    plan = fr.plan(texts, speeds)
    outs = g.synth_list(plan, prm)
    for u in range(plan.n_utts):
        want, _ = orc.synth(prm, plan.utt_ops(u), float(speeds[u]))
        if len(want) != len(outs[u]) or not np.array_equal(want, outs[u]):
            bad += 1
            print("MISMATCH variant", variant, "utt", u, "speed", float(speeds[u]), "len", len(want), len(outs[u]), repr(texts[u][:60]))
    print(f"variant {variant}: {plan.n_utts} utterances checked, mismatches so far {bad}, {time.time() - t0:.0f} s", flush=True)
print("STRESS", "FAILED" if bad else "OK", bad)
sys.exit(1 if bad else 0)
