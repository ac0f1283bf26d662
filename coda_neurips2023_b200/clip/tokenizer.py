"""Byte-level BPE tokenizer of OpenAI CLIP, for the class-prompt text features.

Behavioural mirror of CLIP/clip/simple_tokenizer.py:62-132 and `tokenize` (CLIP/clip/clip.py:279-319): same
byte -> printable-unicode table, same merge list (lines 1 .. 48894 of `bpe_simple_vocab_16e6.txt.gz`), same
vocabulary order (256 byte symbols, their `</w>` forms, the merges, then <|startoftext|> / <|endoftext|>), same
pre-tokenisation pattern, lower-casing and whitespace folding; `tokenize` returns (n, 77) int32 rows
`[sot] + ids + [eot]`, zero padded.  The vocabulary file is NOT part of this package: it is looked up in a CoDA
checkout (`$CODA_CLIP_BPE`, `./CLIP/clip/`, `./clip/`) and a `FileNotFoundError` tells the caller (model_3detr)
that no real text features are available.  `ftfy` is optional here (absent -> text is used as is; the CoDA
class names are plain ASCII).
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Iterable, List, Union

import regex
import torch

VOCAB_FILE = "bpe_simple_vocab_16e6.txt.gz"
N_MERGES = 49152 - 256 - 2
SOT, EOT = "<|startoftext|>", "<|endoftext|>"
_SPLIT = regex.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                       regex.IGNORECASE)


def find_vocab() -> str:
    cands = [os.environ.get("CODA_CLIP_BPE", ""), os.path.join("CLIP", "clip", VOCAB_FILE),
             os.path.join("clip", VOCAB_FILE), VOCAB_FILE]
    for c in cands:
        if c and os.path.isfile(c):
            return c
    raise FileNotFoundError(f"{VOCAB_FILE} not found (set CODA_CLIP_BPE or run from a CoDA checkout)")


def _byte_symbols() -> List[str]:
    """symbol of byte b: the printable latin-1 characters stand for themselves, the other 68 bytes are mapped,
    in increasing byte order, to the code points 256, 257, ..."""
    keep = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, spill = [], 0
    for b in range(256):
        if b in keep:
            table.append(chr(b))
        else:
            table.append(chr(256 + spill))
            spill += 1
    return table


class ByteBPE:
    def __init__(self, vocab_path: str | None = None):
        path = vocab_path or find_vocab()
        with gzip.open(path, "rt", encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges = [tuple(ln.split()) for ln in lines[1:N_MERGES + 1]]
        self.byte_sym = _byte_symbols()
        # vocabulary order of the reference: symbols in ITS table order (kept bytes first, then the remapped ones)
        kept = [b for b in range(256) if ord(self.byte_sym[b]) < 256]
        moved = [b for b in range(256) if ord(self.byte_sym[b]) >= 256]
        base = [self.byte_sym[b] for b in kept + moved]
        vocab = base + [s + "</w>" for s in base] + ["".join(m) for m in merges] + [SOT, EOT]
        self.ids = {tok: i for i, tok in enumerate(vocab)}
        self.rank = {m: i for i, m in enumerate(merges)}
        self._memo = {SOT: [self.ids[SOT]], EOT: [self.ids[EOT]]}

    def _merge_word(self, symbols: List[str]) -> List[str]:
        """greedy BPE: repeatedly fuse every occurrence of the adjacent pair with the lowest merge rank"""
        while len(symbols) > 1:
            best, best_rank = None, None
            for pair in zip(symbols, symbols[1:]):
                r = self.rank.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            out, i = [], 0
            while i < len(symbols):
                if i + 1 < len(symbols) and symbols[i] == best[0] and symbols[i + 1] == best[1]:
                    out.append(best[0] + best[1])
                    i += 2
                else:
                    out.append(symbols[i])
                    i += 1
            symbols = out
        return symbols

    def _word_ids(self, word: str) -> List[int]:
        hit = self._memo.get(word)
        if hit is None:
            syms = [self.byte_sym[b] for b in word.encode("utf-8")]
            syms[-1] += "</w>"
            hit = self._memo[word] = [self.ids[s] for s in self._merge_word(syms)]
        return hit

    def encode(self, text: str) -> List[int]:
        try:
            import ftfy
            text = ftfy.fix_text(text)
        except ImportError:
            pass
        text = html.unescape(html.unescape(text)).strip()
        text = regex.sub(r"\s+", " ", text).strip().lower()
        out: List[int] = []
        for word in _SPLIT.findall(text):
            out.extend(self._word_ids(word))
        return out


@lru_cache(maxsize=2)
def _default(path: str | None = None) -> ByteBPE:
    return ByteBPE(path)


def tokenize(texts: Union[str, Iterable[str]], context_length: int = 77, truncate: bool = False,
             vocab_path: str | None = None) -> torch.Tensor:
    """(n, context_length) int32: [sot] + ids + [eot], zero padded (CLIP/clip/clip.py:279-319)."""
    if isinstance(texts, str):
        texts = [texts]
    bpe = _default(vocab_path)
    sot, eot = bpe.ids[SOT], bpe.ids[EOT]
    rows = torch.zeros((len(list(texts)) if not isinstance(texts, list) else len(texts), context_length), dtype=torch.int32)
    for i, t in enumerate(texts):
        ids = [sot] + bpe.encode(t) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = eot
        rows[i, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
    return rows
