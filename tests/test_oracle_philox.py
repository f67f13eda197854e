"""Known-answer tests for the Philox4x32-10 restatement (Random123 kat_vectors)."""
import numpy as np

from oracle import philox_oracle as po


def test_philox_known_answers():
    # philox4x32 10 <ctr> <key> -> <out>, Random123 examples/kat_vectors
    r = po.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = po.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(x) for x in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = po.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(x) for x in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_discrete_codes_distribution():
    codes = po.discrete_codes(7, np.arange(2000), 500, (1 / 6, 1 / 6, 2 / 3))
    frac = np.bincount(codes.ravel(), minlength=3) / codes.size
    assert np.allclose(frac, [1 / 6, 1 / 6, 2 / 3], atol=3e-3)
