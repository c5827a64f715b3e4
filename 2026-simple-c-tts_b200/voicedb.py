"""voice.db reader/writer and the seeded synthetic voice used by tests and bench.

The on-disk format is the reference's (ctts.h:84-111 header/index structs,
writer ctts.c:1000-1080, reader ctts.c:1117-1165).  `build_voice_db` produces
the same bytes `ctts build` produces for the same units (checked against the
compiled reference in tests/test_voicedb.py), so the GPU box can create the
database without the reference tree.

Synthetic voice: 38 letters + CV/CCV/V(C) syllables,
harmonic-pulse audio, PCM16 mono 22050 Hz, with small DC offsets, noise,
different f0 per unit and quiet lead-in / tail segments so that DC removal,
RMS gain, pitch smoothing and silence trimming all fire.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

CTTS_MAGIC = 0x53545443  # ctts.h:22
CTTS_VERSION = 1
SAMPLE_RATE = 22050

HEADER_FMT = "<12I16s"  # ctts.h:84-98, 64 bytes
INDEX_DTYPE = np.dtype(
    [  # ctts.h:101-111, 32 bytes
        ("hash", "<u4"),
        ("string_offset", "<u4"),
        ("string_len", "<u2"),
        ("char_count", "<u2"),
        ("audio_offset", "<u4"),
        ("sample_count", "<u4"),
        ("flags", "<u4"),
        ("next_hash", "<u4"),
        ("reserved", "<u4"),
    ]
)
assert INDEX_DTYPE.itemsize == 32 and struct.calcsize(HEADER_FMT) == 64


def fnv1a(data: bytes) -> int:
    """ctts_hash, ctts.c:224."""
    h = 2166136261
    for b in data:
        h ^= b
        h = (h * 16777619) & 0xFFFFFFFF
    return h


def normalize_text(text: str) -> str:
    """ctts_normalize, ctts.c:271 (ASCII + the four accented capitals of :238-246)."""
    out = []
    for ch in text:
        cp = ord(ch)
        if 0x41 <= cp <= 0x5A:
            cp += 32
        elif cp == 0xC9:
            cp = 0xE9
        elif cp == 0xD3:
            cp = 0xF3
        elif cp == 0xD4:
            cp = 0xF4
        elif cp == 0xC7:
            cp = 0xE7
        out.append(chr(cp))
    return "".join(out)


def build_voice_db(units: list[tuple[str, np.ndarray]]) -> bytes:
    """Serialise (text, int16 samples) units exactly as ctts_build_database does.

    Order: char_count descending then strcmp of the UTF-8 bytes (compare_units,
    ctts.c:931-937).  Texts must be unique after normalisation (qsort is not
    stable; duplicates would make the order implementation-defined).
    """
    recs = []
    for text, pcm in units:
        t = normalize_text(text).encode("utf-8")
        recs.append((t, len(t.decode("utf-8")), np.ascontiguousarray(pcm, dtype="<i2")))
    if len({r[0] for r in recs}) != len(recs):
        raise ValueError("duplicate unit texts")
    recs.sort(key=lambda r: (-r[1], r[0]))
    n = len(recs)
    strings_size = sum(len(r[0]) + 1 for r in recs)
    total_samples = sum(len(r[2]) for r in recs)
    max_chars = max((r[1] for r in recs), default=0)
    hts = 1
    while hts < n / 0.7:  # ctts.c:989-991
        hts *= 2
    index_offset = 64
    hash_table_offset = index_offset + n * 32
    strings_offset = hash_table_offset + hts * 4
    audio_offset = strings_offset + strings_size
    header = struct.pack(
        HEADER_FMT,
        CTTS_MAGIC,
        CTTS_VERSION,
        n,
        SAMPLE_RATE,
        16,
        index_offset,
        strings_offset,
        audio_offset,
        total_samples,
        max_chars,
        hts,
        hash_table_offset,
        b"\0" * 16,
    )
    index = np.zeros(n, dtype=INDEX_DTYPE)
    table = np.full(hts, 0xFFFFFFFF, dtype="<u4")
    spos = 0
    apos = 0
    for i, (t, cc, pcm) in enumerate(recs):
        h = fnv1a(t)
        index[i] = (h, spos, len(t), cc, apos, len(pcm), 0, 0xFFFFFFFF, 0)
        slot = h % hts
        if table[slot] == 0xFFFFFFFF:
            table[slot] = i
        else:
            prev = int(table[slot])
            while index[prev]["next_hash"] != 0xFFFFFFFF:
                prev = int(index[prev]["next_hash"])
            index[prev]["next_hash"] = i
        spos += len(t) + 1
        apos += len(pcm)
    strings = b"".join(r[0] + b"\0" for r in recs)
    pcm_all = np.concatenate([r[2] for r in recs]) if recs else np.zeros(0, "<i2")
    return header + index.tobytes() + table.tobytes() + strings + pcm_all.tobytes()


@dataclass
class VoiceDB:
    """Parsed view of a voice.db byte string (the layout ctts_init maps, ctts.c:1144-1159)."""

    raw: bytes
    unit_count: int
    max_unit_chars: int
    hash_table_size: int
    index: np.ndarray
    hash_table: np.ndarray
    strings_offset: int
    audio_offset: int
    total_samples: int

    @property
    def pcm(self) -> np.ndarray:
        return np.frombuffer(self.raw, dtype="<i2", count=self.total_samples, offset=self.audio_offset) \
            if self.audio_offset % 2 == 0 else \
            np.frombuffer(self.raw[self.audio_offset:self.audio_offset + 2 * self.total_samples], dtype="<i2")

    def unit_text(self, i: int) -> str:
        e = self.index[i]
        o = self.strings_offset + int(e["string_offset"])
        return self.raw[o:o + int(e["string_len"])].decode("utf-8")

    def unit_pcm(self, i: int) -> np.ndarray:
        e = self.index[i]
        a = int(e["audio_offset"])
        return self.pcm[a:a + int(e["sample_count"])]


def parse_voice_db(raw: bytes) -> VoiceDB:
    (magic, version, n, sr, bits, index_off, strings_off, audio_off, total, max_chars, hts,
     ht_off, _res) = struct.unpack_from(HEADER_FMT, raw, 0)
    if magic != CTTS_MAGIC or version != CTTS_VERSION:
        raise ValueError("not a CTTS voice.db (magic/version)")
    index = np.frombuffer(raw, dtype=INDEX_DTYPE, count=n, offset=index_off)
    table = np.frombuffer(raw, dtype="<u4", count=hts, offset=ht_off)
    return VoiceDB(raw, n, max_chars, hts, index, table, strings_off, audio_off, total)


# --------------------------------------------------------------------------
# synthetic voice
# --------------------------------------------------------------------------

LETTERS = list("abcdefghijklmnopqrstuvwxyz") + list("áàâãéêíóôõúç")

_ONSETS = ["", "b", "c", "d", "f", "g", "j", "l", "m", "n", "p", "r", "s", "t", "v", "x", "z",
           "ch", "lh", "nh", "qu", "gu", "br", "cr", "dr", "fr", "gr", "pr", "tr", "vr",
           "bl", "cl", "fl", "gl", "pl", "rr", "ss", "ç", "h"]
_NUCLEI = ["a", "e", "i", "o", "u", "á", "â", "ã", "é", "ê", "í", "ó", "ô", "õ", "ú",
           "ai", "ei", "oi", "ui", "au", "eu", "ou", "ão", "õe", "ãe", "ia", "io", "ua"]
_CODAS = ["", "s", "r", "l", "m", "n", "z"]


def synthetic_inventory(n_syllables: int = 1749, seed: int = 2026) -> tuple[list[str], list[str]]:
    """Letters + a seeded choice of syllable texts (all distinct, none equal to a letter)."""
    rng = np.random.default_rng(seed)
    core = []  # every open syllable: needed so ordinary words are coverable
    for o in _ONSETS:
        for v in _NUCLEI:
            core.append(o + v)
    closed = []
    for o in _ONSETS:
        for v in _NUCLEI[:15]:
            for c in _CODAS[1:]:
                closed.append(o + v + c)
    seen = set(LETTERS)
    syll = []
    for s in core:
        if s not in seen:
            seen.add(s)
            syll.append(s)
    order = rng.permutation(len(closed))
    for k in order:
        if len(syll) >= n_syllables:
            break
        s = closed[int(k)]
        if s not in seen:
            seen.add(s)
            syll.append(s)
    return list(LETTERS), syll[:n_syllables]


def synth_unit_pcm(rng: np.random.Generator) -> np.ndarray:
    """One harmonic-pulse unit: f0 in [100,180) Hz, 8 harmonics 1/h, 120-350 ms."""
    dur_ms = rng.uniform(120.0, 350.0)
    n = int(dur_ms * SAMPLE_RATE / 1000.0)
    f0 = rng.uniform(100.0, 180.0)
    peak = rng.uniform(3000.0, 12000.0)
    dc = rng.uniform(-150.0, 150.0)
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    # slow vibrato keeps autocorrelation peaks from being exactly periodic
    drift = 1.0 + 0.01 * np.sin(2 * np.pi * rng.uniform(2.0, 6.0) * t + rng.uniform(0, 6.28))
    phase = 2 * np.pi * f0 * np.cumsum(drift) / SAMPLE_RATE
    x = np.zeros(n)
    for h in range(1, 9):
        x += np.sin(h * phase + rng.uniform(0, 2 * np.pi)) / h
    x /= np.max(np.abs(x)) + 1e-12
    env = np.ones(n)
    a = int(rng.uniform(5.0, 15.0) * SAMPLE_RATE / 1000.0)
    r = int(rng.uniform(5.0, 15.0) * SAMPLE_RATE / 1000.0)
    env[:a] = 0.5 - 0.5 * np.cos(np.pi * np.arange(a) / a)
    env[n - r:] = 0.5 + 0.5 * np.cos(np.pi * np.arange(r) / r)
    # quiet lead-in / tail / interior gap (recorded syllables have them)
    u = rng.uniform()
    if u < 0.25:
        q = int(rng.uniform(15.0, 60.0) * SAMPLE_RATE / 1000.0)
        env[:q] = 0.0
        env[q:q + a] = 0.5 - 0.5 * np.cos(np.pi * np.arange(a) / a)
    elif u < 0.5:
        q = int(rng.uniform(15.0, 60.0) * SAMPLE_RATE / 1000.0)
        env[n - q:] = 0.0
        env[n - q - r:n - q] = 0.5 + 0.5 * np.cos(np.pi * np.arange(r) / r)
    elif u < 0.6:
        q = int(rng.uniform(20.0, 50.0) * SAMPLE_RATE / 1000.0)
        s = int(rng.uniform(0.3, 0.6) * n)
        env[s:s + q] = 0.0
    x = peak * x * env + dc + rng.normal(0.0, 20.0, n)
    return np.clip(np.round(x), -32768, 32767).astype("<i2")


def synthetic_units(n_syllables: int = 1749, seed: int = 2026) -> tuple[list, list]:
    """(letters, syllables) as lists of (text, pcm)."""
    letters, sylls = synthetic_inventory(n_syllables, seed)
    rng = np.random.default_rng(seed + 1)
    lu = [(t, synth_unit_pcm(rng)) for t in letters]
    su = [(t, synth_unit_pcm(rng)) for t in sylls]
    return lu, su


def synthetic_voice_db(n_syllables: int = 1749, seed: int = 2026) -> bytes:
    lu, su = synthetic_units(n_syllables, seed)
    return build_voice_db(lu + su)


def write_wav(path: str, pcm: np.ndarray) -> None:
    """PCM16 mono 22050 Hz, 44-byte header (same container as ctts_write_wav, ctts.c:809)."""
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    data = pcm.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, SAMPLE_RATE, SAMPLE_RATE * 2, 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)


def write_dataset(root: str, letters: list, syllables: list) -> None:
    """Dataset directory in the layout `ctts build` reads (ctts.c:3956-3959, index lines :877-884)."""
    for sub, idx_name, units in (("letters", "letters.txt", letters),
                                 ("syllables", "sillabes.txt", syllables)):
        wav_dir = os.path.join(root, sub, "wavs")
        os.makedirs(wav_dir, exist_ok=True)
        lines = []
        for k, (text, pcm) in enumerate(units):
            name = f"{sub[0]}{k:05d}"
            write_wav(os.path.join(wav_dir, name + ".wav"), pcm)
            lines.append(f"{name}|{text}|{text}\n")
        with open(os.path.join(root, sub, idx_name), "w", encoding="utf-8") as f:
            f.writelines(lines)
