"""Single-utterance latency through the drop-in call (BASELINE configs[0] / [1]: 'olá mundo' at 1.0 and 1.5),
next to the compiled reference on one host core."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import harness as H
gpu = importlib.import_module("2026-simple-c-tts_b200.gpu")
db = H.synthetic_db()
fr = H.front.Front(db, H.shipped_config(), H.NORM_CSV)
prm = fr.params()
g = gpu.GpuSynth(db, 0)
ref = None
if H.have_reference():
    import tempfile
    p = os.path.join(tempfile.mkdtemp(), "voice.db"); open(p, "wb").write(db)
    ref = H.Reference(p)
for text, speed in (("olá mundo", 1.0), ("olá mundo", 1.5), (H.corpus.batch(1, seed=1)[0], 1.0), (H.corpus.batch(1, seed=1)[0], 1.5)):
    plan = fr.plan([text], [speed])
    off = g.layout(plan)
    import torch
    host = torch.empty(int(off[-1]), dtype=torch.int16).pin_memory().numpy()
    for _ in range(5): g.synth_batch(plan, prm, host, off)
    ts = []
    for _ in range(30):
        t = time.perf_counter(); _, _, cnt = g.synth_batch(plan, prm, host, off); ts.append(time.perf_counter() - t)
    tp = []
    for _ in range(30):
        t = time.perf_counter(); fr.plan([text], [speed]); tp.append(time.perf_counter() - t)
    line = f"{len(text):4d} chars speed {speed}: {int(cnt[0])/22050:.2f} s audio; gpu synth_batch median {1e3*np.median(ts):.3f} ms (min {1e3*min(ts):.3f}); plan {1e3*np.median(tp):.3f} ms"
    if ref:
        tr = []
        for _ in range(10):
            t = time.perf_counter(); ref.synth(text, speed); tr.append(time.perf_counter() - t)
        line += f"; reference ctts_synthesize median {1e3*np.median(tr):.3f} ms"
    print(line)
