"""Known-answer test of the numpy Philox emulation (Random123 kat_vectors, philox4x32 10 rounds)."""
import numpy as np

from philox_ref import philox4x32, u01_halfopen, u01_open


def _run(ctr, key):
    k = key[0] | (key[1] << 32)
    return [int(x) for x in philox4x32(k, *ctr)]


def test_random123_known_answers():
    assert _run((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _run((f, f, f, f), (f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _run((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniform_ranges():
    z, f = np.uint32(0), np.uint32(0xffffffff)
    assert u01_halfopen(z, z) == 0.0 and u01_halfopen(f, f) < 1.0
    assert 0.0 < u01_open(z, z) and u01_open(f, f) < 1.0
