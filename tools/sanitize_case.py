"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): a few utterances, both speeds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
import numpy as np
import harness as H
gpu = importlib.import_module("2026-simple-c-tts_b200.gpu")
db = H.small_db()
fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
prm = fr.params()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
texts = H.corpus.batch(n, seed=5, target_chars=60) + ["olá mundo", "a, b. c? d!"]
speeds = [1.0] * n + [1.5, 1.0]
plan = fr.plan(texts, speeds)
g = gpu.GpuSynth(db, 0)
outs = g.synth_list(plan, prm)
orc = H.Oracle(db)
for u, got in enumerate(outs):
    want, _ = orc.synth(prm, plan.utt_ops(u), speeds[u])
    assert np.array_equal(got, want), u
print("sanitize case ok", [len(o) for o in outs])
