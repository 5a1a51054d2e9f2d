"""The reference's OWN model classes (src/models/gat.py, src/models/tgn.py, imported unmodified from /root/reference)
against the committed fixture and against the drop-in layer -- CPU side.

* With ``torch_geometric.nn.GATConv`` stubbed to the CPU oracle layer, the reference's classes reproduce
  ``tests/golden/reference_models_golden.npz`` (so the fixture the GPU tests use IS the reference wrapper code's output)
  and agree bit for bit with the oracle's restated wrappers (``OracleGAT`` / ``OracleTemporalGNN``).
* With the stub pointing at THIS repo's ``GATConv``, the reference's classes construct, expose the same state_dict keys
  and load the reference's checkpoints with ``strict=True`` -- the one-line import swap of INTEGRATION.md.
Skipped where /root/reference is absent (the GPU box)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "models")),
                                reason="/root/reference is only present in the authoring container")


def _gen():
    spec = importlib.util.spec_from_file_location("make_reference_goldens", os.path.join(HERE, "golden", "make_reference_goldens.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_fixture_is_the_reference_wrappers_output():
    gen = _gen()
    fresh = gen.reference_outputs()
    gold = np.load(os.path.join(HERE, "golden", "reference_models_golden.npz"))
    assert sorted(fresh) == sorted(gold.files)
    for k in gold.files:
        assert np.array_equal(np.asarray(fresh[k]), gold[k]), k
    # ... and the oracle's restated wrappers (what golden_small.npz was made with) are the same function
    small = np.load(os.path.join(HERE, "golden", "golden_small.npz"))
    assert np.array_equal(gold["gat_logits_f32"], small["gat_logits"])
    assert np.array_equal(gold["tgn_logits_f32"], small["tgn_logits"]) and np.array_equal(gold["tgn_hidden_f32"], small["tgn_hidden"])


def test_reference_classes_accept_the_drop_in_layer():
    from gnn_fraud_detection_b200 import GATConv
    gen = _gen()
    RefGAT, RefTGN = gen.load_reference_models(GATConv)
    for cls, ck in ((RefGAT, "gat_model.pt"), (RefTGN, "tgn_model.pt")):
        sd = torch.load(os.path.join(REF, "results", ck), map_location="cpu", weights_only=False)
        m = cls(165, 64, 1, num_layers=3)
        assert all(type(l) is GATConv for l in m.gat_layers)
        assert sorted(m.state_dict().keys()) == sorted(sd.keys())
        m.load_state_dict(sd, strict=True)
        for l in m.gat_layers:                      # the reference's ctor call: heads=8, concat=False, dropout=p
            assert (l.heads, l.concat, l.dropout) == (8, False, 0.2) and l.lin_dst is l.lin_src
        with pytest.raises(RuntimeError):           # the layer has no CPU path: the model must be moved to the GPU
            m(torch.randn(4, 165), torch.zeros(2, 0, dtype=torch.long))
