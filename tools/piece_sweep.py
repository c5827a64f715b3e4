"""text -> host PCM (ctts_b200_synth_texts) for several piece sizes: piece_sweep.py speed1|mixed [sizes...]"""
import sys, os, time, importlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, torch, harness as H
pkg = importlib.import_module("2026-simple-c-tts_b200"); gpu = importlib.import_module("2026-simple-c-tts_b200.gpu"); pipe = importlib.import_module("2026-simple-c-tts_b200.pipeline")
db = H.synthetic_db(); fr = pkg.front.Front(db, H.shipped_config(), H.NORM_CSV)
n = 4096
mixed = sys.argv[1] == "mixed"
sizes = [int(a) for a in sys.argv[2:]] or [128, 192, 256, 384]
texts = pkg.corpus.batch(n, seed=1234); speeds = pkg.corpus.mixed_speeds(n, seed=99) if mixed else np.ones(n, np.float32)
g = gpu.GpuSynth(db, 0)
plan = fr.plan(texts, speeds)
host = torch.empty(int(g.layout(plan)[-1]) + 4096, dtype=torch.int16).pin_memory().numpy()
tb = pipe.TextBatch(texts, speeds)
for size in sizes:
    pipe.synth_texts(fr, g, tb, host, piece_utts=size)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); off, cnt, used, tm = pipe.synth_texts(fr, g, tb, host, piece_utts=size); ts.append(time.perf_counter() - t0)
    print(f"{sys.argv[1]} pieces of {size}: {1e3 * min(ts):.1f} / {1e3 * sorted(ts)[len(ts) // 2]:.1f} ms (min / median), {tm.pieces} pieces, {used * 2 / min(ts) / 1e9:.1f} GB/s")
