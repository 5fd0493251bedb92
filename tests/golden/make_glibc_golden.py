"""Generates tests/golden/glibc_rand42.json straight from glibc (srand(42); rand()) -- the generator the reference's
initWeight (optimize-gcn/gcn.h:838-852) uses.  Independent of oracle/ and of the engine."""
import ctypes
import json
import math
import os

libc = ctypes.CDLL("libc.so.6")
libc.srand(42)
raw = [libc.rand() for _ in range(16)]
libc.srand(42)
limit = math.sqrt(6.0 / (2 + 3))
w = [float(libc.rand()) / 2147483647.0 * 2 * limit - limit for _ in range(6)]
json.dump({"rand_after_srand42": raw, "initWeight_2x3": w},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "glibc_rand42.json"), "w"), indent=1)
print(raw[:4])
