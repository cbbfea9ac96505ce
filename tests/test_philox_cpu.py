"""Known-answer test of the numpy Philox4x32-10 restatement (tests/philox_ref.py) that the GPU tests use to rebuild the
dropout keep-masks of the fused kernels (csrc/philox.cuh).  Vectors: the kat_vectors of Random123 (Salmon et al., SC'11),
philox4x32 with 10 rounds: counter (4 words), key (2 words) -> 4 words."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import philox_ref  # noqa: E402

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = philox_ref.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want, (ctr, key, [hex(int(x)) for x in got])


def test_keep_mask_rate_and_determinism():
    m1 = philox_ref.keep_mask(0x1234_5678_9ABC_DEF0, 1, 4, 64, 256, 0.1)
    m2 = philox_ref.keep_mask(0x1234_5678_9ABC_DEF0, 1, 4, 64, 256, 0.1)
    assert m1.shape == (1, 4, 64, 256) and (m1 == m2).all()
    assert abs(m1.mean() - 0.9) < 0.01                           # 65 536 Bernoulli(0.9) draws: sigma = 0.0012
    assert not (m1 == philox_ref.keep_mask(0x1234_5678_9ABC_DEF1, 1, 4, 64, 256, 0.1)).all()
    assert philox_ref.keep_mask(7, 1, 4, 8, 64, 0.0).all()
