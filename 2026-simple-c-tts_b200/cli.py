"""`ctts synth`-compatible driver over the B200 back end (row 12 of SURVEY.md section 8a: the device emits
PCM, the 44-byte RIFF header of ctts_write_wav, ctts.c:809, is host work).

  python -m 2026-simple-c-tts_b200.cli synth <voice.db> "text" <out.wav> [speed]      # ctts.c:3970-4027
  python -m 2026-simple-c-tts_b200.cli synth-batch <voice.db> <texts.tsv> <out_dir>   # "speed<TAB>text" lines

Like the reference CLI it reads config.yaml and normalization.csv from the working directory when they
exist (ctts.c:3636, :3990), clamps the speed to [0.5, 2.0] (ctts.c:3976-3981) and uses default_speed when no
speed is given and the config sets one (ctts.c:3993).  There is no CPU fallback: without a CUDA device it fails.
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import front as _front
from . import voicedb as _voicedb


def _open(db_path: str):
    from . import gpu as _gpu
    with open(db_path, "rb") as f:
        db = f.read()
    cfg = _front.load_config("config.yaml" if os.path.exists("config.yaml") else None)
    norm = "normalization.csv" if os.path.exists("normalization.csv") else None
    fr = _front.Front(db, cfg, norm)
    return fr, _gpu.GpuSynth(db, int(os.environ.get("CTTS_GPU_DEVICE", "0"))), cfg


def _clamp(speed: float) -> float:
    return min(2.0, max(0.5, speed))


def synth(db_path: str, text: str, out_wav: str, speed: float | None = None) -> int:
    fr, g, cfg = _open(db_path)
    if speed is None:
        speed = cfg.default_speed if cfg.default_speed != 1.0 else 1.0
    plan = fr.plan([text], [np.float32(_clamp(float(speed)))])
    pcm = g.synth_list(plan, fr.params())[0]
    _voicedb.write_wav(out_wav, pcm)
    print(f"Synthesized {len(pcm)} samples ({len(pcm) / _voicedb.SAMPLE_RATE:.2f} s), "
          f"units found {int(plan.found[0])}, missing {int(plan.missing[0])}")
    return 0


def synth_batch(db_path: str, tsv: str, out_dir: str) -> int:
    fr, g, _ = _open(db_path)
    speeds, texts = [], []
    with open(tsv, encoding="utf-8") as f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            s, _, t = line.partition("\t")
            speeds.append(_clamp(float(s)))
            texts.append(t)
    plan = fr.plan(texts, np.asarray(speeds, dtype=np.float32))
    os.makedirs(out_dir, exist_ok=True)
    for u, pcm in enumerate(g.synth_list(plan, fr.params())):
        _voicedb.write_wav(os.path.join(out_dir, f"{u:06d}.wav"), pcm)
    print(f"Synthesized {len(texts)} utterances into {out_dir}")
    return 0


def main(argv: list[str]) -> int:
    if len(argv) >= 4 and argv[0] == "synth":
        return synth(argv[1], argv[2], argv[3], float(argv[4]) if len(argv) > 4 else None)
    if len(argv) == 4 and argv[0] == "synth-batch":
        return synth_batch(argv[1], argv[2], argv[3])
    print(__doc__)
    return 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
