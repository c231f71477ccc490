"""Transcribes the reference's own known-answer vectors for the hot path into
tests/golden/reference_vectors.json.  Every entry cites the reference test it
comes from (paths relative to /root/reference/test/librawspeed).  The expected
values are the literals / generator formulae of those tests, restated here; the
script does NOT run the oracle -- it only writes what the reference asserts.

Run:  python tests/golden/make_golden.py
"""
import json
import os

LSB, MSB, MSB16, MSB32, JPEG = 0, 1, 2, 3, 4


def gen_ones_le(zeros_to_output, zeros_outputted):
    # bitstreams/BitStreamerTest.h:205-222 (GenOnesLE)
    v, bits, curr = [], 0, -1
    for _ in range(29):
        if zeros_to_output == zeros_outputted:
            bits |= 1 << curr
            zeros_to_output += 1
            zeros_outputted = 0
        v.append(bits & 0xFFFFFFFF)
        zeros_outputted += 1
        curr += 1
    return v


def gen_ones_be(zeros_to_output, zeros_outputted):
    # bitstreams/BitStreamerTest.h:223-238 (GenOnesBE)
    v, bits = [], 0
    for _ in range(29):
        if zeros_to_output == zeros_outputted:
            bits |= 1
            zeros_to_output += 1
            zeros_outputted = 0
        v.append(bits & 0xFFFFFFFF)
        zeros_outputted += 1
        bits = (bits << 1) & 0xFFFFFFFFFF
    return v


def pad8(b):
    return list(b) + [0] * (8 - len(b))


# per-pump byte patterns: BitStreamerLSBTest.cpp:35-50, BitSteramerMSBTest.cpp:35-50,
# BitStreamerMSB16Test.cpp:35-50, BitStreamerMSB32Test.cpp:35-50, BitStreamerJPEGTest.cpp:45-71
ONES = {
    LSB: [0b01001011, 0b10000100, 0b00100000, 0b11110000],
    MSB: [0b10100100, 0b01000010, 0b00001000, 0b00011111],
    MSB16: [0b01000010, 0b10100100, 0b00011111, 0b00001000],
    MSB32: [0b00011111, 0b00001000, 0b01000010, 0b10100100],
    JPEG: [0b10100100, 0b01000010, 0b00001000, 0b00011111],
}
INVONES = {
    LSB: [0b00100101, 0b01000010, 0b00010000, 0b11111000],
    MSB: [0b11010010, 0b00100001, 0b00000100, 0b00001111],
    MSB16: [0b00100001, 0b11010010, 0b00001111, 0b00000100],
    MSB32: [0b00001111, 0b00000100, 0b00100001, 0b11010010],
    JPEG: [0b11010010, 0b00100001, 0b00000100, 0b00001111],
}


def pump_vectors():
    out = []
    for order in (LSB, MSB, MSB16, MSB32, JPEG):
        le = order == LSB
        sat = [0xFF, 0, 0xFF, 0, 0xFF, 0, 0xFF, 0] if order == JPEG else pad8([0xFF] * 4)
        pats = {
            # Pattern<*, ZerosTag> BitStreamerTest.h:163-169
            "zeros": (pad8([]), [0] * 8, [0] * 29),
            # OnesTag: element(i)=1, data = GenOnesLE(0,-1) / GenOnesBE(1,0)
            "ones": (pad8(ONES[order]), [1] * 8, gen_ones_le(0, -1) if le else gen_ones_be(1, 0)),
            # InvOnesTag: element(i)=1<<(i-1), data = GenOnesLE(1,0) / GenOnesBE(0,-1)
            "invones": (pad8(INVONES[order]), [0] + [1 << (i - 1) for i in range(1, 8)],
                        gen_ones_le(1, 0) if le else gen_ones_be(0, -1)),
            # SaturatedTag BitStreamerTest.h:190-199
            "saturated": (sat, [(1 << i) - 1 for i in range(8)],
                          [(1 << i) - 1 for i in range(29)]),
        }
        for name, (data, element, datafn) in pats.items():
            out.append({
                "order": order, "pattern": name, "bytes": data,
                # GetTest / PeekTest (:132-168): getBits(len)==element(len), len=1..7
                "get_lens": list(range(1, 8)), "get_expect": element[1:8],
                # IncreasingPeekLengthTest (:170-186): fresh peekBits(len)==data(len), len=1..28
                "peek_expect": datafn[1:29],
                "cite": "bitstreams/BitStreamerTest.h:132-255",
            })
    return out


def main():
    doc = {
        "pumps": pump_vectors(),
        # BitStreamerJPEGTest.cpp:73-85 (FF00 is a data FF)
        "jpeg_ff00": {"bytes": [0xFF, 0x00, 0b10100100, 0b01000010, 0b00001000, 0b00011111, 0, 0, 0, 0],
                      "lens": [8, 1, 2, 3, 4, 5, 6, 7], "expect": [0xFF, 1, 1, 1, 1, 1, 1, 1],
                      "cite": "bitstreams/BitStreamerJPEGTest.cpp:73-85"},
        # BitStreamerJPEGTest.cpp:87-101 (FFxx ends the stream: >= 96 zero bits)
        "jpeg_end_marker": {"ends": list(range(1, 0xFF)), "tail": [0xFF, 0xFF, 0xFF, 0xFF, 0, 0, 0, 0],
                            "nbits": 96, "cite": "bitstreams/BitStreamerJPEGTest.cpp:87-101"},
        # codes/HuffmanCodeTest.cpp:502-580 (extend truth table)
        "extend": ([[0, l, -((1 << l) - 1)] for l in range(1, 17)] +
                   [[(1 << l) - 1, l, (1 << l) - 1] for l in range(1, 17)] +
                   [[1 << l, l, 1] for l in range(1, 17)] +
                   [[0b00, 1, -1], [0b01, 1, 1], [0b10, 1, 1], [0b11, 1, 3],
                    [0b00, 2, -3], [0b01, 2, -2], [0b10, 2, 2], [0b11, 2, 3],
                    [0b00, 3, -7], [0b01, 3, -6], [0b10, 3, -5], [0b11, 3, -4]]),
        # codes/HuffmanCodeTest.cpp:600-633 (canonical code assignment)
        "code_symbols": [
            {"ncpl": [1], "symbols": [[0b0, 1]]},
            {"ncpl": [0, 1], "symbols": [[0b00, 2]]},
            {"ncpl": [0, 2], "symbols": [[0b00, 2], [0b01, 2]]},
            {"ncpl": [0, 3], "symbols": [[0b00, 2], [0b01, 2], [0b10, 2]]},
            {"ncpl": [1, 1], "symbols": [[0b0, 1], [0b10, 2]]},
            {"ncpl": [1, 2], "symbols": [[0b0, 1], [0b10, 2], [0b11, 2]]},
        ],
        # codes/HuffmanCodeTest.cpp:366-401 (count validation)
        "ncpl_validation": [
            {"ncpl": [], "ok": False}, {"ncpl": [0], "ok": False}, {"ncpl": [0, 0], "ok": False},
            {"ncpl": [0, 0, 0, 0, 0, 0, 0, 162], "ok": True},
            {"ncpl": [0, 0, 0, 0, 0, 0, 0, 163], "ok": False},
            {"ncpl": [1], "ok": True}, {"ncpl": [2], "ok": True}, {"ncpl": [3], "ok": False},
            {"ncpl": [1, 2], "ok": True}, {"ncpl": [1, 3], "ok": False},
            {"ncpl": [2, 1], "ok": False}, {"ncpl": [0, 4], "ok": True},
            {"ncpl": [0, 5], "ok": False},
        ],
        # codes/HuffmanTableTest.cpp:69-132 (decode known answers, MSB pump)
        "huff_decode": [
            {"ncpl": [2], "values": [4, 8], "full": False,
             "bytes": [0b01010101] * 4, "n": 32, "expect": [4, 8] * 16,
             "cite": "codes/HuffmanTableTest.cpp:69-85"},
            {"ncpl": [2], "values": [7, 15], "full": True,
             "bytes": [0b00000000, 0b11010101, 0b01010101, 0b01111111], "n": 3,
             "expect": [-127, 21845, 127], "cite": "codes/HuffmanTableTest.cpp:87-102"},
            {"ncpl": [1], "values": [4], "full": False, "bytes": [0b01000000, 0, 0, 0], "n": 2,
             "expect": [4, "RDE"], "cite": "codes/HuffmanTableTest.cpp:104-117"},
            {"ncpl": [1], "values": [1], "full": True, "bytes": [0b00100000, 0, 0, 0], "n": 2,
             "expect": [-1, "RDE"], "cite": "codes/HuffmanTableTest.cpp:119-132"},
        ],
        # SURVEY appendix A.1 examples (cross-checked against the compiled reference)
        "unpack_examples": [
            {"order": MSB, "bytes": [0x3c, 0x5e, 0x81], "bps": 14, "first": 3863},
            {"order": MSB, "bytes": [0x3c, 0x5e, 0x81], "bps": 12, "first": 965, "second": 3713},
            {"order": LSB, "bytes": [0x3c, 0x5e, 0x81], "bps": 14, "first": 7740},
            {"order": LSB, "bytes": [0x3c, 0x5e, 0x81], "bps": 12, "first": 3644, "second": 2069},
            {"order": MSB16, "bytes": [0x3c, 0x5e], "bps": 14, "first": 6031},
            {"order": MSB32, "bytes": [0x3c, 0x5e, 0x81, 0xb4], "bps": 14, "first": 11552},
        ],
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
