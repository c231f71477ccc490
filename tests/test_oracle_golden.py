"""Pins the oracle against the reference's own unit-test vectors
(tests/golden/reference_vectors.json, transcribed by tests/golden/make_golden.py
from /root/reference/test/librawspeed).  CPU only."""
import json
import os

import pytest

from oracle import port

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


@pytest.mark.parametrize("v", G["pumps"], ids=lambda v: "%d-%s" % (v["order"], v["pattern"]))
def test_pump_patterns(v):
    data = bytes(v["bytes"])
    assert port.pump_getbits(v["order"], data, v["get_lens"]) == v["get_expect"]
    for ln, want in zip(range(1, 29), v["peek_expect"]):
        assert port.pump_getbits(v["order"], data, [ln]) == [want], ln


def test_jpeg_ff00_is_ff():
    v = G["jpeg_ff00"]
    assert port.pump_getbits(port.JPEG, bytes(v["bytes"]), v["lens"]) == v["expect"]


def test_jpeg_ffxx_is_the_end():
    v = G["jpeg_end_marker"]
    for end in v["ends"]:
        data = bytes([0xFF, end] + v["tail"])
        assert port.pump_getbits(port.JPEG, data, [1] * v["nbits"]) == [0] * v["nbits"]


def test_extend_truth_table():
    for diff, ln, want in G["extend"]:
        assert port.huff_extend(diff, ln) == want


def test_canonical_code_generation():
    for v in G["code_symbols"]:
        ncpl = v["ncpl"] + [0] * (16 - len(v["ncpl"]))
        h = port.Huff(ncpl, [0] * sum(ncpl))
        assert [list(s) for s in h.symbols()] == v["symbols"]


def test_ncpl_validation():
    for v in G["ncpl_validation"]:
        ncpl = v["ncpl"] + [0] * (16 - len(v["ncpl"]))
        if v["ok"]:
            port.Huff(ncpl, [0] * sum(ncpl), full=False)
        else:
            with pytest.raises(port.RawDecoderException):
                port.Huff(ncpl, [0] * min(sum(ncpl), 200), full=False)


def test_huffman_decode_known_answers():
    for v in G["huff_decode"]:
        ncpl = v["ncpl"] + [0] * (16 - len(v["ncpl"]))
        h = port.Huff(ncpl, v["values"], full=v["full"])
        data = bytes(v["bytes"])
        if "RDE" in v["expect"]:
            k = v["expect"].index("RDE")
            assert h.decode(data, k, order=port.MSB) == v["expect"][:k]
            with pytest.raises(port.RawDecoderException):
                h.decode(data, v["n"], order=port.MSB)
        else:
            assert h.decode(data, v["n"], order=port.MSB) == v["expect"]


def test_unpack_examples():
    for v in G["unpack_examples"]:
        data = bytes(v["bytes"]) + bytes(8)
        lens = [v["bps"]] * (2 if "second" in v else 1)
        got = port.pump_getbits(v["order"], data, lens)
        assert got[0] == v["first"]
        if "second" in v:
            assert got[1] == v["second"]
