#!/usr/bin/env python
"""Generates tests/golden/hts_golden.json from the REFERENCE's own objects (oracle/_ref/libhts_ref.so, built by
oracle/Makefile from /root/reference/src/htscodecs, unmodified): for every genozip order byte (codec_htscodecs.c:17-20),
13 stream kinds and a set of sizes incl. the container edge cases, the length, the SHA-256 and the first 24 bytes of the
reference's compressed output of a seeded input (tests/datagen.py).  The fixtures travel; /root/reference does not.

    python tests/golden/make_golden.py        (needs oracle/_ref)
"""
import hashlib, json, os, sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc
from datagen import stream, KINDS

CODECS = [("rans", "RANB"), ("rans", "RANW"), ("rans", "RANb"), ("rans", "RANw"),
          ("arith", "ARTB"), ("arith", "ARTW"), ("arith", "ARTb"), ("arith", "ARTw")]
SIZES = [0, 1, 7, 8, 20, 21, 49, 50, 257, 4097, 70001]


def cases():
    for kind, name in CODECS:
        for dk in KINDS:
            for n in SIZES:
                yield kind, name, dk, n, 1000 + n


def main():
    out = []
    for kind, name, dk, n, seed in cases():
        c = orc.compress("ref", kind, stream(dk, n, seed), orc.ORDER[name])
        out.append(dict(codec=name, kind=dk, n=n, seed=seed, len=int(c.size), sha256=hashlib.sha256(c.tobytes()).hexdigest(), head=c[:24].tobytes().hex()))
    json.dump(dict(source="oracle/_ref/libhts_ref.so = /root/reference/src/htscodecs/{rANS_static4x16pr,arith_dynamic,pack,rle}.c, reference flags",
                   generator="tests/golden/make_golden.py", cases=out), open(os.path.join(HERE, "hts_golden.json"), "w"), indent=0)
    print(len(out), "cases")


if __name__ == "__main__":
    main()
