"""Seeded synthetic Portuguese text batches for tests and bench.

Sentences are drawn from a small vocabulary that exercises what the reference's
acceptance sentences exercise (generate_samples.sh sections: questions,
exclamations, commas, numbers, abbreviations, hiatus, initial R, intervocalic S):
all four phrase types, digits (number expansion), words the normalisation rules
rewrite, hyphens and the punctuation pauses.
"""
from __future__ import annotations

import numpy as np

WORDS = """
a o e de da do em um uma que não sim com para por mais mas como quando onde muito pouco
casa mesa rosa coisa música brasil mundo olá bom dia boa tarde noite hoje amanhã ontem
tempo vida ano mês semana hora minuto cidade país rua praia sol lua mar rio terra céu
água fogo vento chuva flor árvore livro carta palavra língua pessoa homem mulher criança
amigo família trabalho escola carro trem avião navio porta janela cadeira comida café
pão leite fruta maçã banana laranja rato rei rua rádio roupa rápido carro terra
falar dizer fazer ver ir vir ter ser estar poder querer saber ficar passar chegar
gosto gosta gostamos falou disse fez viu foi veio tem era está pode quer sabe ficou
grande pequeno novo velho bonito feio alto baixo forte fraco feliz triste calmo
azul verde vermelho branco preto claro escuro quente frio doce amargo
praia areia ideia meio feio saia maio joia apoio cheio veia
chave chuva filho folha ninho vinho queijo quilo guerra guia pronto branco triste
flor claro globo plano frio livre
presidente telefone computador universidade importante diferente necessário possível
""".split()

ABBREV = ["Dr.", "Sr.", "Sra.", "Prof.", "km", "kg", "etc."]
ENDINGS = [".", ".", ".", "?", "?", "!", ",", ";", ""]
MID = [",", ",", ",", ";", ":", " -", ""]


ONSETS = "b c d f g j l m n p r s t v z ch lh nh br cr dr fr gr pr tr bl cl fl pl qu gu".split() + [""]
NUCLEI = "a e i o u á é í ó ú ã õ ai ei oi ui au eu ou ão".split()
CODAS = ["", "", "", "", "s", "r", "m", "l", "n"]


class Vocabulary:
    """`size` word types with Zipf-Mandelbrot frequencies (p ~ 1 / (rank + 2.7)): the words above first (shuffled),
    then pronounceable pseudo-words of 1-4 syllables.  For the vocabulary-size sensitivity of anything that gains
    from repeated words (bench.py --vocab)."""

    def __init__(self, size: int, seed: int = 4321):
        rng = np.random.default_rng(seed)
        seen = dict.fromkeys(WORDS)
        words = list(seen)
        rng.shuffle(words)
        while len(words) < size:
            w = ""
            for _ in range(int(rng.integers(1, 5))):
                w += ONSETS[int(rng.integers(len(ONSETS)))] + NUCLEI[int(rng.integers(len(NUCLEI)))]
            w += CODAS[int(rng.integers(len(CODAS)))]
            if w not in seen:
                seen[w] = None
                words.append(w)
        self.words = words[:size]
        p = 1.0 / (np.arange(len(self.words)) + 2.7)
        self.cdf = np.cumsum(p / p.sum())

    def draw(self, rng: np.random.Generator) -> str:
        return self.words[min(int(np.searchsorted(self.cdf, rng.uniform())), len(self.words) - 1)]


def sentence(rng: np.random.Generator, target_chars: int = 200, vocab: Vocabulary | None = None) -> str:
    def word() -> str:
        return vocab.draw(rng) if vocab is not None else WORDS[int(rng.integers(len(WORDS)))]

    parts: list[str] = []
    length = 0
    first = True
    while length < target_chars:
        u = rng.uniform()
        if u < 0.05:
            w = str(int(rng.integers(0, 3000)))
        elif u < 0.08:
            w = str(int(rng.integers(1000, 2000000)))
        elif u < 0.11:
            w = ABBREV[int(rng.integers(len(ABBREV)))]
        elif u < 0.13:
            w = word() + "-" + word()
        else:
            w = word()
        if first:
            w = w[0].upper() + w[1:]
            first = False
        v = rng.uniform()
        if v < 0.12 and length + len(w) < target_chars - 10:
            w += MID[int(rng.integers(len(MID)))]
        elif v < 0.16 and length + len(w) < target_chars - 10:
            w += [".", "?", "!"][int(rng.integers(3))]
            first = True
        parts.append(w)
        length += len(w) + 1
    text = " ".join(parts)
    text = text.rstrip(",;: -")
    return text + ENDINGS[int(rng.integers(len(ENDINGS)))]


def batch(n: int, seed: int = 1234, target_chars: int = 200, vocab: Vocabulary | None = None) -> list[str]:
    rng = np.random.default_rng(seed)
    return [sentence(rng, target_chars, vocab) for _ in range(n)]


def mixed_speeds(n: int, seed: int = 99) -> np.ndarray:
    """Speeds uniform in {0.5 .. 2.0 step 0.1} excluding 1.0 (BASELINE.json configs[3])."""
    rng = np.random.default_rng(seed)
    grid = np.array([s / 10.0 for s in range(5, 21) if s != 10], dtype=np.float32)
    return grid[rng.integers(0, len(grid), size=n)]
