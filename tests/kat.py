"""Shared helpers: run the Appendix-B known-answer cases through any encoder back-end.

A back-end is anything with the oracle's Python surface (oracle.Restated, oracle.Reference,
or the GPU adapter in tests/gpu_backend.py): bs_encode_batch / spu_encode / xa_encode.
"""
import json
import os

import numpy as np

from psxavenc_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FNV_OFFSET = 1469598103934665603
FNV_PRIME = 1099511628211
M64 = (1 << 64) - 1


def load_kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def fnv(data):
    return synth.fnv1a64(np.ascontiguousarray(data).tobytes())


def frames_hash(out, max_size):
    h = FNV_OFFSET
    for i in range(out.shape[0]):
        h ^= fnv(out[i, :max_size])
        h = (h * FNV_PRIME) & M64
    return h


def run_bs_case(backend, case, fdct):
    frames = synth.gen_frames(0, 8, case["w"], case["h"], case["noise"])
    out, res = backend.bs_encode_batch(case["codec"], case["w"], case["h"], frames, case["max"], fdct)
    return "%016x" % frames_hash(out, case["max"]), res


def xa_input(case):
    pcm = synth.gen_pcm(case["n"], case["ch"], case["seed"])
    return np.concatenate([pcm, np.zeros((4032, case["ch"]), np.int16)])
