"""B200-native batched audio-assembly back end for CTTS (jonathandasilvasantos/2026-simple-c-tts).

Only the hot path behind `ctts synth` lives here: the host front end that turns
text into a CSR batch plan (front.py -> csrc/front), the CUDA executor behind the
C-ABI `ctts_gpu_synth_batch()` (gpu.py -> csrc/gpu), and the voice.db / corpus
helpers tests and bench use.  The directory name is not a Python identifier;
import it with importlib.import_module("2026-simple-c-tts_b200").
"""
from . import _build, corpus, front, hostgather, sharding, voicedb  # noqa: F401  (gpu / pipeline load CUDA libraries: import them explicitly)

__all__ = ["_build", "corpus", "front", "hostgather", "sharding", "voicedb"]
