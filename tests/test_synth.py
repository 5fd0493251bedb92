"""Synthetic inputs of the named shapes (tools/synth.py): exact sizes, symmetric edge lists without self loops,
the `vid % T` partition of data_transform.py:25, exactly 10 % inter-party edges for BASELINE configs[2], and the reference's
three text formats (graph_io_util.h:66-147, kernel_harness.h:37-44) round-tripping."""
import numpy as np
import pytest

from tools import synth


@pytest.mark.parametrize("shape,T", [("cora", 2), ("pubmed", 2), ("citeseer", 4), ("cora_small", 2)])
def test_named_shapes_have_the_survey_sizes(shape, T):
    N, E, F, H, C, _ = synth.SHAPES[shape]
    g = synth.make(shape, T)
    e = g["edges"]
    assert e.shape == (E, 2) and g["N"] == N and g["feats"].shape == (N, F) and g["labels"].shape == (N,)
    assert (e[:, 0] != e[:, 1]).all(), "no self loops"
    fwd = set(map(tuple, e.tolist()))
    assert len(fwd) == E and all((d, s) in fwd for s, d in fwd), "every undirected pair is stored twice, once per direction"
    assert np.array_equal(g["tid"], np.arange(N) % T)
    assert g["cfg"]["input_dim"] == F and g["cfg"]["hidden_dim"] == H and g["cfg"]["num_labels"] == C
    assert 0 <= g["labels"].min() and g["labels"].max() < C
    again = synth.make(shape, T)
    assert np.array_equal(again["edges"], e) and np.array_equal(again["feats"], g["feats"]), "seeded: same graph every time"


def test_block_partition_with_ten_percent_inter_party_edges():
    g = synth.make("citeseer", 4, 0.1)  # BASELINE configs[2]
    E = g["E"]
    assert g["inter_party_edges"] == 2 * int(round(E // 2 * 0.1))
    tid = g["tid"]
    assert (np.diff(tid) >= 0).all() and set(tid.tolist()) == {0, 1, 2, 3}, "contiguous blocks"


def test_reference_file_formats_round_trip(tmp_path):
    g = synth.make("cora_small", 2)
    prefix = str(tmp_path / "cora_small")
    synth.write_reference_files(g, prefix)
    edges = np.loadtxt(prefix + ".edge.preprocessed", dtype=np.int64).reshape(-1, 2)
    part = np.loadtxt(prefix + ".part.preprocessed", dtype=np.int64).reshape(-1, 2)
    vert = np.loadtxt(prefix + ".vertex.preprocessed").reshape(g["N"], -1)
    assert np.array_equal(edges, g["edges"])
    assert np.array_equal(part[:, 0], np.arange(g["N"])) and np.array_equal(part[:, 1], g["tid"])
    assert np.array_equal(vert[:, 0], np.arange(g["N"]))
    assert np.array_equal(vert[:, 1:-1], g["feats"]) and np.array_equal(vert[:, -1].astype(np.int32), g["labels"])
