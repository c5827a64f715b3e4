#!/usr/bin/env python
"""The literal `./ctts synth` figure SURVEY.md 8(d) asks for once: one PROCESS per utterance of the stock reference
command line (oracle/_ref/ctts, built from /root/reference/ctts.c by oracle/Makefile), as many at a time as the box
has cores -- exec, mmap, the 49 regcomp calls, synthesis and the WAV write all inside the timed region (ctts.c:3970-4030).
TEST / BASELINE infrastructure.   usage: spawn_baseline.py [n_utts] > profiles/rNN_spawn_baseline.json"""
import json, os, shutil, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cores = os.cpu_count() or 1
texts = H.corpus.batch(n, seed=1234, target_chars=200)
with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, "voice.db"), "wb").write(H.synthetic_db())
    shutil.copy(H.SHIPPED_YAML, os.path.join(d, "config.yaml"))
    shutil.copy(H.NORM_CSV, os.path.join(d, "normalization.csv"))
    open(os.path.join(d, "duration_rules.csv"), "w").close()

    # xargs spawns the processes (a fork from this Python process would cost more than the synthesis itself)
    def run(idx):
        args = b"".join(f"{i}".encode() + b"\0" + texts[i].encode("utf-8") + b"\0" for i in idx)
        cmd = ["xargs", "-0", "-n", "2", "-P", str(cores), "sh", "-c",
               f'exec "{H.REF_CLI}" synth voice.db "$2" "o$1.wav" 1.0 > /dev/null 2>&1', "sh"]
        return subprocess.run(cmd, input=args, cwd=d).returncode

    run(range(min(cores, n)))                            # warm the page cache
    t0 = time.perf_counter()
    rc_all = run(range(n))
    dt = time.perf_counter() - t0
    res = [(rc_all, max(os.path.getsize(os.path.join(d, f"o{i}.wav")) - 44, 0) // 2) for i in range(n)]
assert all(rc == 0 for rc, _ in res), "ctts synth failed"
samples = sum(s for _, s in res)
print(json.dumps({"what": "stock `ctts synth` spawned once per utterance, one process per core at a time (exec + mmap + regcomp + synth + WAV write timed)",
                  "utterances": n, "cores": cores, "seconds": dt, "audio_seconds": samples / 22050,
                  "audio_seconds_per_second": samples / 22050 / dt, "ms_per_utterance_per_core": 1e3 * dt * cores / n}))
