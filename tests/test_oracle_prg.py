"""Pins the oracle's PRG (oracle/cgb_oracle.c orc_chacha20_block / orc_prg_fill) against RFC 8439 and OpenSSL."""
import json
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
RFC_KEY = [0x03020100, 0x07060504, 0x0B0A0908, 0x0F0E0D0C, 0x13121110, 0x17161514, 0x1B1A1918, 0x1F1E1D1C]


def test_rfc8439_block_vector(oracle):
    # RFC 8439 section 2.3.2: key 00..1f, nonce 00:00:00:09:00:00:00:4a:00:00:00:00, block counter 1
    out = oracle.chacha20_block(RFC_KEY, 1, [0x09000000, 0x4A000000, 0x00000000])
    expect = [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
              0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    assert [int(v) for v in out] == expect


def test_stream_layout_against_openssl_golden(oracle):
    cases = json.load(open(os.path.join(HERE, "golden", "chacha20_openssl.json")))["cases"]
    for c in cases:
        words = np.array([int(w) for w in c["words"]], dtype=np.uint64)
        got = oracle.prg_fill(c["key"], c["stream"], c["word_offset"], words.size)
        assert np.array_equal(got, words)
        # unaligned windows into the same stream
        got = oracle.prg_fill(c["key"], c["stream"], c["word_offset"] + 3, words.size - 5)
        assert np.array_equal(got, words[3:-2])


def test_against_cryptography_live(oracle):
    cryptography = pytest.importorskip("cryptography")
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms

    key = [11, 22, 33, 44, 55, 66, 77, 88]
    stream = 0xABCDEF0123
    nonce = struct.pack("<4I", 5, stream & 0xFFFFFFFF, stream >> 32, 0)
    enc = Cipher(algorithms.ChaCha20(struct.pack("<8I", *key), nonce), mode=None).encryptor()
    ks = enc.update(b"\x00" * 64 * 6)
    words = np.frombuffer(ks, dtype="<u8")
    assert np.array_equal(oracle.prg_fill(key, stream, 5 * 8, words.size), words)


def test_empty_and_single_word(oracle):
    assert oracle.prg_fill(RFC_KEY, 1, 0, 0).size == 0
    a = oracle.prg_fill(RFC_KEY, 1, 0, 16)
    for w in range(16):
        assert oracle.prg_fill(RFC_KEY, 1, w, 1)[0] == a[w]
