"""Regenerates / cross-checks tests/golden/kat.json from the UNMODIFIED reference sources.

  python tests/golden/make_kat.py            # recompute every vector with oracle/_ref and compare
  python tests/golden/make_kat.py --write    # rewrite kat.json with the recomputed values

Needs /root/reference (to build oracle/_ref/libpsxav_ref.so through oracle/Makefile) or a prebuilt
oracle/_ref. Inputs are the integer generators of psxavenc_b200/synth.py;
each case keeps its parameters, only the hashes / lengths / states are recomputed.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle                      # noqa: E402
from psxavenc_b200 import synth    # noqa: E402
from tests import kat              # noqa: E402


def recompute(ref, data):
    out = {"_comment": data["_comment"], "bs": [], "spu": [], "xa": []}
    for case in data["bs"]:
        c = dict(case)
        for name, fdct in (("sse2", oracle.FDCT_SSE2), ("islow", oracle.FDCT_ISLOW)):
            h, res = kat.run_bs_case(ref, case, fdct)
            c[name] = h
            c["last_" + name] = [int(res[-1, 0]), int(res[-1, 1])]
            qs = set(int(q) for q in res[:, 2])
            assert len(qs) == 1, "case expects one quant scale for all 8 frames"
            c["q"] = qs.pop()
        out["bs"].append(c)
    for case in data["spu"]:
        c = dict(case)
        pcm = synth.gen_pcm(case["n"], case["ch"], case["seed"])
        states = [oracle.ChannelState() for _ in range(case["ch"])]
        enc = np.concatenate([ref.spu_encode(states[ch], pcm, case["count"], case["ch"], offset=ch)
                              for ch in range(case["ch"])])
        c["len"] = int(len(enc))
        c["hash"] = "%016x" % kat.fnv(enc)
        if "prev1" in case:
            c["prev1"], c["prev2"] = int(states[0].prev1), int(states[0].prev2)
            c["first_block"] = enc[:16].tobytes().hex()
        out["spu"].append(c)
    for case in data["xa"]:
        c = dict(case)
        st = oracle.new_states()
        enc = ref.xa_encode(case["format"], case["ch"] == 2, 37800, case["bits"], 1, 2, st, kat.xa_input(case),
                            case["n"], 7, finalize=True)
        c["len"] = int(len(enc))
        c["hash"] = "%016x" % kat.fnv(enc)
        out["xa"].append(c)
    return out


def main():
    path = os.path.join(kat.GOLDEN, "kat.json")
    with open(path) as f:
        data = json.load(f)
    oracle.build()
    fresh = recompute(oracle.Reference(), data)
    if "--write" in sys.argv:
        with open(path, "w") as f:
            json.dump(fresh, f, indent=1)
            f.write("\n")
        print("rewrote", path)
        return 0
    same = fresh == data
    print("kat.json %s the unmodified reference (%d BS x 2 FDCTs, %d SPU, %d XA vectors)"
          % ("matches" if same else "DIFFERS from", len(data["bs"]), len(data["spu"]), len(data["xa"])))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
