"""Runs the C++ shim demo (tests/shim_demo.cpp): the reference's primitive API names bound onto the C ABI, driven by an
ALICE and a BOB thread with the reference's call sequence and aliasing."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_shim_demo_runs_reference_call_sequence():
    from cognn_b200.host import build as hb

    exe = hb.build_shim_demo()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SHIM_DEMO_OK" in r.stdout, r.stdout + r.stderr


def test_shim_header_compiles_standalone():
    from cognn_b200.host import build as hb

    assert os.path.exists(hb.build_shim_demo())
