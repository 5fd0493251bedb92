"""The drop-in claim on the GPU: the reference's own harness.cpp + engine + GCN operator headers, compiled unchanged against
cognn_b200/host/shim/include (oracle/_ref, built by oracle/build_ref.py where /root/reference exists), run one process per party
on the CUDA library and reproduce the epoch oracle's loss / accuracy log.  See tests/test_reference_dropin.py."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _need():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gcn-optimize")):
        pytest.skip("oracle/_ref/gcn-optimize was not shipped to this box")


@pytest.mark.timeout(900)
def test_reference_code_runs_on_the_cuda_library_and_matches_oracle():
    from tests.graphs import small_graph
    from tests import refdrop
    from tests.test_reference_dropin import CFG, check_against_oracle, run_until_complete

    _need()
    g = small_graph(n=60, n_edges=220, F=10, C=4, T=2, seed=5)
    gd = refdrop.graph_dict(g["edges"], g["tid"], g["feats"], g["labels"], CFG)
    out = run_until_complete("gcn-optimize", gd, 2, 12, 2, False, 33000)
    check_against_oracle(out, g, 2, 12, 1e-3)


@pytest.mark.timeout(900)
def test_reference_inference_and_original_operators_on_the_cuda_library():
    from oracle import epoch as oep
    from tests.graphs import small_graph
    from tests import refdrop
    from tests.test_reference_dropin import CFG, check_against_oracle, run_until_complete

    _need()
    g = small_graph(n=60, n_edges=220, F=10, C=4, T=2, seed=6)
    gd = refdrop.graph_dict(g["edges"], g["tid"], g["feats"], g["labels"], CFG)
    out = run_until_complete("gcn-inference-optimize", gd, 2, 2, 1, False, 33300)
    check_against_oracle(out, g, 2, 2, 1e-3)
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gcn-original")):
        out = run_until_complete("gcn-original", gd, 2, 4, 1, False, 33600)
        o = oep.EpochOracle(g["edges"], g["tid"], 2, g["feats"], g["labels"], CFG)
        o.run(6)
        for p in range(2):
            want = [m["loss"] for m in o.log if m["party"] == p]
            assert abs(out[p]["loss"][0] - want[0]) < 2e-2
