"""ctypes binding of the host front end (include/ctts_front.h): text -> batch plan."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _build

# ctts_plan_op, include/ctts_plan.h (32 bytes)
OP_DTYPE = np.dtype(
    [("kind", "<u2"), ("flags", "<u2"), ("a", "<u4"), ("b", "<u4"),
     ("f0", "<f4"), ("f1", "<f4"), ("f2", "<f4"), ("e0", "<f4"), ("e1", "<f4")]
)
assert OP_DTYPE.itemsize == 32

OP_UNIT, OP_SILENCE, OP_FADE_OUT, OP_WORD_END, OP_MARK = 1, 2, 3, 4, 5
UNIT_AFTER_BOUNDARY = 1
WE_TRIM, WE_INTON, WE_CIRCUMFLEX, WE_ENERGY = 1, 2, 4, 8


class Config(C.Structure):
    """ctts_front_config == CTTSConfig (ctts.h:44-77)."""

    _fields_ = [
        ("crossfade_ms", C.c_float), ("crossfade_vowel_ms", C.c_float),
        ("crossfade_s_ending_ms", C.c_float), ("crossfade_r_ending_ms", C.c_float),
        ("vowel_to_consonant_factor", C.c_float), ("word_pause_ms", C.c_float),
        ("unknown_silence_ms", C.c_float), ("fade_in_ms", C.c_float), ("fade_out_ms", C.c_float),
        ("remove_word_silence", C.c_int), ("silence_threshold", C.c_float),
        ("min_silence_ms", C.c_float), ("remove_dc_offset", C.c_int),
        ("normalize_level", C.c_float), ("compression", C.c_float), ("default_speed", C.c_float),
        ("min_speed", C.c_float), ("max_speed", C.c_float), ("max_pitch_change", C.c_float),
        ("print_units", C.c_int), ("print_timing", C.c_int),
    ]


class AssemblyParams(C.Structure):
    """ctts_assembly_params, include/ctts_plan.h."""

    _fields_ = [
        ("fade_in_samples", C.c_uint32), ("min_silence_samples", C.c_uint32),
        ("silence_threshold", C.c_float), ("target_rms", C.c_float),
        ("remove_dc_offset", C.c_uint32), ("reserved", C.c_uint32 * 3),
    ]


class CBatchPlan(C.Structure):
    """ctts_batch_plan, include/ctts_plan.h."""

    _fields_ = [
        ("n_utts", C.c_uint32), ("n_ops", C.c_uint32),
        ("utt_op_begin", C.POINTER(C.c_uint32)), ("speed", C.POINTER(C.c_float)),
        ("ops", C.c_void_p),
    ]


@dataclass
class BatchPlan:
    """Host copy of a CSR batch plan (numpy-owned)."""

    utt_op_begin: np.ndarray  # uint32 [n_utts+1]
    speed: np.ndarray         # float32 [n_utts]
    ops: np.ndarray           # OP_DTYPE [n_ops]
    found: np.ndarray | None = None    # units_found per utterance
    missing: np.ndarray | None = None  # units_missing per utterance

    @property
    def n_utts(self) -> int:
        return int(self.speed.shape[0])

    def as_c(self) -> CBatchPlan:
        p = CBatchPlan()
        p.n_utts = self.n_utts
        p.n_ops = int(self.ops.shape[0])
        p.utt_op_begin = self.utt_op_begin.ctypes.data_as(C.POINTER(C.c_uint32))
        p.speed = self.speed.ctypes.data_as(C.POINTER(C.c_float))
        p.ops = self.ops.ctypes.data
        return p

    def utt_ops(self, u: int) -> np.ndarray:
        return self.ops[int(self.utt_op_begin[u]):int(self.utt_op_begin[u + 1])]

    def select(self, idx) -> "BatchPlan":
        """Sub-plan holding utterances `idx` (used to shard a batch across ranks)."""
        idx = np.asarray(idx, dtype=np.int64)
        chunks = [self.utt_ops(int(u)) for u in idx]
        begin = np.zeros(len(idx) + 1, dtype=np.uint32)
        if len(idx):
            begin[1:] = np.cumsum([len(c) for c in chunks])
        ops = np.concatenate(chunks) if chunks else np.zeros(0, OP_DTYPE)
        return BatchPlan(begin, np.ascontiguousarray(self.speed[idx]), np.ascontiguousarray(ops),
                         None if self.found is None else self.found[idx],
                         None if self.missing is None else self.missing[idx])


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build_front())
        L.ctts_front_config_defaults.argtypes = [C.POINTER(Config)]
        L.ctts_front_config_load.argtypes = [C.POINTER(Config), C.c_char_p]
        L.ctts_front_open.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t,
                                      C.POINTER(Config), C.c_char_p]
        L.ctts_front_close.argtypes = [C.c_void_p]
        L.ctts_front_rule_count.argtypes = [C.c_void_p]
        L.ctts_front_rule_count.restype = C.c_uint32
        L.ctts_front_unit_count.argtypes = [C.c_void_p]
        L.ctts_front_unit_count.restype = C.c_uint32
        L.ctts_front_params.argtypes = [C.c_void_p, C.POINTER(AssemblyParams)]
        L.ctts_front_normalize_text.argtypes = [C.c_void_p, C.c_char_p]
        L.ctts_front_normalize_text.restype = C.c_void_p
        L.ctts_front_free.argtypes = [C.c_void_p]
        L.ctts_front_plan_batch.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float),
                                            C.c_uint32, C.POINTER(CBatchPlan), C.POINTER(C.c_uint32)]
        L.ctts_front_plan_free.argtypes = [C.POINTER(CBatchPlan)]
        L.ctts_front_word_end_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ctts_front_plan_bounds.argtypes = [C.c_void_p, C.POINTER(CBatchPlan), C.c_void_p,
                                             C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def load_config(path: str | None) -> Config:
    cfg = Config()
    lib().ctts_front_config_load(C.byref(cfg), path.encode() if path else None)
    return cfg


class Front:
    """Host front end over voice.db bytes; mirrors `ctts_init` + the text half of `ctts_synthesize`."""

    def __init__(self, voice_db: bytes, config: Config | None = None, normalization_csv: str | None = None):
        self._db = voice_db  # keep alive: borrowed by the C side
        self._buf = (C.c_char * len(voice_db)).from_buffer_copy(voice_db)
        self.config = config if config is not None else load_config(None)
        h = C.c_void_p()
        rc = lib().ctts_front_open(C.byref(h), C.addressof(self._buf), len(voice_db), C.byref(self.config),
                                   normalization_csv.encode() if normalization_csv else None)
        if rc != 0:
            raise RuntimeError(f"ctts_front_open failed: {rc}")
        self._h = h

    def close(self) -> None:
        if self._h:
            lib().ctts_front_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rule_count(self) -> int:
        return int(lib().ctts_front_rule_count(self._h))

    def params(self) -> AssemblyParams:
        p = AssemblyParams()
        lib().ctts_front_params(self._h, C.byref(p))
        return p

    def normalize_text(self, text: str) -> str:
        p = lib().ctts_front_normalize_text(self._h, text.encode("utf-8"))
        try:
            return C.string_at(p).decode("utf-8", errors="replace")
        finally:
            lib().ctts_front_free(p)

    def plan(self, texts: list[str], speeds=None) -> BatchPlan:
        n = len(texts)
        arr = (C.c_char_p * max(n, 1))(*[t.encode("utf-8") for t in texts])
        sp = None
        if speeds is not None:
            sp_np = np.ascontiguousarray(speeds, dtype=np.float32)
            assert sp_np.shape == (n,)
            sp = sp_np.ctypes.data_as(C.POINTER(C.c_float))
        stats = np.zeros(2 * max(n, 1), dtype=np.uint32)
        cp = CBatchPlan()
        rc = lib().ctts_front_plan_batch(self._h, arr, sp, n, C.byref(cp),
                                         stats.ctypes.data_as(C.POINTER(C.c_uint32)))
        if rc != 0:
            raise RuntimeError(f"ctts_front_plan_batch failed: {rc}")
        try:
            begin = np.ctypeslib.as_array(cp.utt_op_begin, shape=(n + 1,)).copy()
            speed = np.ctypeslib.as_array(cp.speed, shape=(max(n, 1),))[:n].copy()
            n_ops = int(cp.n_ops)
            raw = C.string_at(cp.ops, n_ops * OP_DTYPE.itemsize) if n_ops else b""
            ops = np.frombuffer(raw, dtype=OP_DTYPE).copy()
        finally:
            lib().ctts_front_plan_free(C.byref(cp))
        st = stats[:2 * n].reshape(n, 2)
        return BatchPlan(begin, speed, ops, st[:, 0].copy(), st[:, 1].copy())

    def word_end_op(self, phrase_type: int, word_index: int, total_words: int) -> np.ndarray:
        """The WORD_END op for (phrase type, word index, word count): scalar half of apply_phrase_intonation."""
        op = np.zeros(1, dtype=OP_DTYPE)
        rc = lib().ctts_front_word_end_op(self._h, phrase_type, word_index, total_words, op.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ctts_front_word_end_op failed: {rc}")
        return op

    def bounds(self, plan: BatchPlan):
        """(pre, out, region) host upper bounds per utterance, see ctts_front_plan_bounds."""
        n = plan.n_utts
        pre = np.zeros(n, dtype=np.uint64)
        out = np.zeros(n, dtype=np.uint64)
        region = np.zeros(n, dtype=np.uint32)
        cp = plan.as_c()
        rc = lib().ctts_front_plan_bounds(self._h, C.byref(cp), pre.ctypes.data, out.ctypes.data,
                                          region.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ctts_front_plan_bounds failed: {rc}")
        return pre, out, region
