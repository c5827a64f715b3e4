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
# repeated words and sentences: canonical regions, tasks resumed at their contour, whole tasks copied
texts += ["bom dia mundo casa", "bom dia mundo casa", "olá bom dia mundo casa", "casa casa casa casa"]
speeds += [1.0, 1.0, 1.0, 0.8]
plan = fr.plan(texts, speeds)
g = gpu.GpuSynth(db, 0)
rp = g.create_plan(plan, prm)
info = rp.info()
print("canonical", info.n_canon_tasks, "resumed or copied", info.n_dedup_tasks, "sources", info.n_source_tasks, "copied whole", info.n_reuse_tasks)
rp.close()
outs = g.synth_list(plan, prm)
orc = H.Oracle(db)
for u, got in enumerate(outs):
    want, _ = orc.synth(prm, plan.utt_ops(u), speeds[u])
    assert np.array_equal(got, want), u
print("sanitize case ok", [len(o) for o in outs])
