"""ctts_gpu_synth_batch vs ctts_gpu_synth_batch_stream on the default bench batch: total time, and when the
first / median chunk reaches the callback (host time since the call started)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import harness as H
pkg = importlib.import_module("2026-simple-c-tts_b200")
gpu = importlib.import_module("2026-simple-c-tts_b200.gpu")
db = H.synthetic_db()
fr = pkg.front.Front(db, H.shipped_config(), H.NORM_CSV)
prm = fr.params()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
plan = fr.plan(pkg.corpus.batch(n, seed=1234, target_chars=200))
g = gpu.GpuSynth(db, 0)
off = g.layout(plan)
host = torch.empty(int(off[-1]), dtype=torch.int16).pin_memory().numpy()
g.synth_batch(plan, prm, host, off)
for k in range(3):
    t = time.perf_counter(); g.synth_batch(plan, prm, host, off); a = time.perf_counter() - t
    stamps = []
    t = time.perf_counter()
    g.synth_batch_stream(plan, prm, lambda p, o, c, b, e: stamps.append((time.perf_counter() - t, e)), host, off)
    b = time.perf_counter() - t
    print(f"plain {1e3*a:.1f} ms; streaming {1e3*b:.1f} ms, {len(stamps)} chunks, first at {1e3*stamps[0][0]:.1f} ms "
          f"({stamps[0][1]} utterances), half of the batch at {1e3*stamps[len(stamps)//2][0]:.1f} ms")
