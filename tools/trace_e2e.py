import sys, os, time, importlib
sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo")
import numpy as np, torch, harness as H
os.environ["CTTS_GPU_TRACE"] = "1"
pkg = importlib.import_module("2026-simple-c-tts_b200"); gpu = importlib.import_module("2026-simple-c-tts_b200.gpu"); pipe = importlib.import_module("2026-simple-c-tts_b200.pipeline")
db = H.synthetic_db(); fr = pkg.front.Front(db, H.shipped_config(), H.NORM_CSV); prm = fr.params()
n = 4096
texts = pkg.corpus.batch(n, seed=1234); speeds = pkg.corpus.mixed_speeds(n, seed=99) if sys.argv[1] == "mixed" else np.ones(n, np.float32)
g = gpu.GpuSynth(db, 0)
plan = fr.plan(texts, speeds)
host = torch.empty(int(g.layout(plan)[-1]) + 4096, dtype=torch.int16).pin_memory().numpy()
tb = pipe.TextBatch(texts, speeds)
sys.stderr.write("WARM\n")
pipe.synth_texts(fr, g, tb, host)
sys.stderr.write("TIMED\n")
t0 = time.perf_counter(); off, cnt, used, tm = pipe.synth_texts(fr, g, tb, host); dt = time.perf_counter() - t0
sys.stderr.write(f"RESULT {1e3*dt:.1f} ms used {used*2/1e9:.2f} GB\n")
