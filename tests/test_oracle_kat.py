"""Pins both oracle back-ends (our restatement and the unmodified reference build) to the
known-answer vectors of SURVEY.md Appendix B — the only golden data that exists for this
path (the reference ships no tests)."""
import numpy as np
import pytest

import oracle
from psxavenc_b200 import synth
from tests import kat

KAT = kat.load_kat()


@pytest.mark.parametrize("case", KAT["bs"], ids=lambda c: "%dx%d-c%d-n%d" % (c["w"], c["h"], c["codec"], c["noise"]))
@pytest.mark.parametrize("fdct", [oracle.FDCT_SSE2, oracle.FDCT_ISLOW], ids=["sse2", "islow"])
def test_bs_kat(any_oracle, case, fdct):
    name = "sse2" if fdct == oracle.FDCT_SSE2 else "islow"
    h, res = kat.run_bs_case(any_oracle, case, fdct)
    assert h == case[name]
    assert (res[:, 2] == case["q"]).all()
    assert list(res[-1, :2]) == case["last_" + name]


@pytest.mark.parametrize("case", KAT["spu"], ids=lambda c: c["name"])
def test_spu_kat(any_oracle, case):
    pcm = synth.gen_pcm(case["n"], case["ch"], case["seed"])
    states = [oracle.ChannelState() for _ in range(case["ch"])]
    out = np.concatenate([any_oracle.spu_encode(states[c], pcm, case["count"], case["ch"], offset=c)
                          for c in range(case["ch"])])
    assert len(out) == case["len"]
    assert "%016x" % kat.fnv(out) == case["hash"]
    if "prev1" in case:
        assert (states[0].prev1, states[0].prev2) == (case["prev1"], case["prev2"])
        assert out[:16].tobytes().hex() == case["first_block"]


@pytest.mark.parametrize("case", KAT["xa"], ids=lambda c: "ch%d-%dbit-f%d" % (c["ch"], c["bits"], c["format"]))
def test_xa_kat(any_oracle, case):
    st = oracle.new_states()
    out = any_oracle.xa_encode(case["format"], case["ch"] == 2, 37800, case["bits"], 1, 2, st, kat.xa_input(case),
                               case["n"], 7, finalize=True)
    assert len(out) == case["len"]
    assert "%016x" % kat.fnv(out) == case["hash"]
