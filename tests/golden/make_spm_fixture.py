#!/usr/bin/env python
"""Generates tests/golden/tiny_spm.model (a small SentencePiece unigram model with byte fallback, a control and a
user-defined symbol) and tests/golden/tiny_spm_decodes.json: id sequences with the text the sentencepiece library
(the C++ library behind the reference's `sentencepiece` crate, batched_asr.rs:667 decode_piece_ids) decodes them
to.  The detokeniser in csrc/spm.cu must reproduce every entry.  No tokenizer .model of the real checkpoints is
reachable offline, hence a trained-here fixture."""
import json
import os
import random
import tempfile

import sentencepiece as spm

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    text = open(os.path.join(ROOT, "SURVEY.md")).read()
    lines = [l.strip() for l in text.split("\n") if len(l.strip()) > 20]
    lines += ["Bonjour à tous, ça va très bien aujourd'hui.", "L'été dernier nous étions à Besançon.", "naïve café déjà vu"] * 20
    with tempfile.TemporaryDirectory() as d:
        corpus = os.path.join(d, "corpus.txt")
        open(corpus, "w").write("\n".join(lines))
        spm.SentencePieceTrainer.train(input=corpus, model_prefix=os.path.join(d, "tiny"), vocab_size=600, model_type="unigram",
                                       byte_fallback=True, character_coverage=0.98, num_threads=1, normalization_rule_name="identity",
                                       user_defined_symbols=["<mark>"], control_symbols=["<ctl>"], minloglevel=2)
        model = open(os.path.join(d, "tiny.model"), "rb").read()
    open(os.path.join(HERE, "tiny_spm.model"), "wb").write(model)
    sp = spm.SentencePieceProcessor(model_proto=model)
    rng = random.Random(0)
    n = sp.get_piece_size()
    cases = []
    for s in ["hello world", "Bonjour à tous", "naïve café", "  two  spaces ", "emoji \U0001F600 ok", "日本語", ""]:
        cases.append(sp.encode(s))
    for _ in range(300):
        k = rng.randint(1, 8)
        cases.append([rng.randrange(n) for _ in range(k)])
    for _ in range(100):      # word-like sequences from the normal part of the vocabulary, as Word messages carry them
        k = rng.randint(1, 5)
        cases.append([rng.randrange(261, n) for _ in range(k)])
    cases += [[0], [1], [2], [3], [4], [0, 0], [3, 300], [300, 3], [5 + 0xE2, 5 + 0x96, 5 + 0x81], [5 + 0xC3], [5 + 0xC3, 5 + 0xA9], [5 + 0xFF, 300]]
    out = [{"ids": c, "text": sp.decode_ids(c)} for c in cases]
    pieces = [{"piece": sp.id_to_piece(i), "score": sp.get_score(i), "control": sp.is_control(i), "unknown": sp.is_unknown(i),
               "byte": sp.is_byte(i)} for i in range(n)]
    json.dump({"vocab_size": n, "decodes": out, "pieces": pieces}, open(os.path.join(HERE, "tiny_spm_decodes.json"), "w"),
              ensure_ascii=True)
    print(n, "pieces,", len(out), "decodes,", len(model), "bytes")


if __name__ == "__main__":
    main()
