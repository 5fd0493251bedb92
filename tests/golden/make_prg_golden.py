"""Generates tests/golden/chacha20_openssl.json: ChaCha20 keystream from the `cryptography` package (OpenSSL),
independent of oracle/ and of the CUDA kernel.  Run once here; the JSON is committed."""
import json
import os
import struct

from cryptography.hazmat.primitives.ciphers import Cipher, algorithms

HERE = os.path.dirname(os.path.abspath(__file__))


def keystream(key_words, stream, first_block, n_blocks):
    key = struct.pack("<8I", *key_words)
    out = b""
    for b in range(first_block, first_block + n_blocks):
        # cryptography's 16-byte nonce = 32-bit LE counter || 96-bit nonce (RFC 8439 layout)
        nonce = struct.pack("<4I", b & 0xFFFFFFFF, stream & 0xFFFFFFFF, (stream >> 32) & 0xFFFFFFFF, (b >> 32) & 0xFFFFFFFF)
        enc = Cipher(algorithms.ChaCha20(key, nonce), mode=None).encryptor()
        out += enc.update(b"\x00" * 64)
    return out


cases = []
for key_words, stream, first_block, n_blocks in [
    ([0x03020100, 0x07060504, 0x0B0A0908, 0x0F0E0D0C, 0x13121110, 0x17161514, 0x1B1A1918, 0x1F1E1D1C], 0, 0, 4),
    ([1, 2, 3, 4, 5, 6, 7, 8], 0x1122334455667788, 0, 3),
    ([0xDEADBEEF, 0, 0xFFFFFFFF, 42, 45, 0x80000000, 7, 9], 45, 0xFFFFFFFE, 4),  # crosses the 32-bit counter wrap
    ([45, 0, 0, 0, 0, 0, 0, 1], (7 << 32) | 3, 1000, 2),
]:
    ks = keystream(key_words, stream, first_block, n_blocks)
    words = list(struct.unpack("<%dQ" % (len(ks) // 8), ks))
    cases.append({"key": key_words, "stream": stream, "word_offset": first_block * 8, "words": [str(w) for w in words]})
json.dump({"source": "cryptography/OpenSSL ChaCha20", "cases": cases}, open(os.path.join(HERE, "chacha20_openssl.json"), "w"), indent=1)
print("wrote", len(cases), "cases")
