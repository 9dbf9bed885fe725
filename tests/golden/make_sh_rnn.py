#!/usr/bin/env python
"""Regenerates tests/golden/sh.rnn from tests/golden/sh.rnnn.

sh.rnnn is the reference's text-format fixture (test_data/sh.rnnn, RNNoise "rnnoise-nu model file version 1").
The transform is the one train/convert_rnnoise.py:18-29 applies: check the header line, split the remainder on
whitespace, take every integer modulo 256 and write the bytes.  Independent of the product's parser
(rnnoise_model_from_text), so tests can compare the two.

    python tests/golden/make_sh_rnn.py            # rewrites sh.rnn next to this script
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def convert(text: str) -> bytes:
    head, _, body = text.partition("\n")
    if head.strip() != "rnnoise-nu model file version 1":
        raise ValueError("unexpected input file format")
    return bytes(int(tok) % 256 for tok in body.split())


if __name__ == "__main__":
    with open(os.path.join(HERE, "sh.rnnn")) as f:
        out = convert(f.read())
    with open(os.path.join(HERE, "sh.rnn"), "wb") as f:
        f.write(out)
    sys.stdout.write("sh.rnn: %d bytes\n" % len(out))
