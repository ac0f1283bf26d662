"""Writes tests/golden/clip_tokens.npz: token rows of the reference's CLIP tokenizer
(/root/reference/CLIP/clip/simple_tokenizer.py + clip.py:tokenize, imported unmodified; `ftfy` is not installed
in this container and is stubbed by the identity, which is exact for these ASCII / latin-1 prompts).

    python tests/golden/make_clip_tokens_golden.py
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/CLIP/clip/simple_tokenizer.py")
CLASSES = ["toilet", "bed", "chair", "bathtub", "sofa", "dresser", "scanner", "fridge", "lamp", "desk", "table",
           "stand", "cabinet", "counter", "bin", "bookshelf", "pillow", "microwave", "sink", "stool", "night stand",
           "tv monitor", "coffee table", "garbage bin", "end table", "dining table", "computer", "whiteboard"]
PROMPTS = ([f"a photo of a {c}" for c in CLASSES] + [f"There is a {c} in the scene." for c in CLASSES[:8]] +
           ["it's the children's bookshelf!!  99 bottles", "naïve café — “quotes” &amp; more", "tv_monitor",
            "A  Photo\tof\nTHE night-stand", "xyzzyqwrt plugh", ""])


def main():
    ftfy = types.ModuleType("ftfy")
    ftfy.fix_text = lambda s: s
    sys.modules["ftfy"] = ftfy
    spec = importlib.util.spec_from_file_location("ref_simple_tokenizer", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    tok = ref.SimpleTokenizer()
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    rows = np.zeros((len(PROMPTS), 77), dtype=np.int32)
    for i, p in enumerate(PROMPTS):
        ids = [sot] + tok.encode(p) + [eot]          # clip.py:302-317
        rows[i, :len(ids)] = ids
    out = Path(__file__).with_name("clip_tokens.npz")
    np.savez_compressed(out, prompts=np.array(PROMPTS), tokens=rows, vocab_size=np.int64(len(tok.encoder)))
    print(out, rows.shape)


if __name__ == "__main__":
    main()
