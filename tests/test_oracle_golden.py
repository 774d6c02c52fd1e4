"""Pins oracle/nets.py to the committed golden vectors (minted by executing the reference's own
net.py files on the paddle shim — tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from tests.util import load_golden, oracle_run

CASES = [("deepfm", "deepfm_d9"), ("deepfm", "deepfm_d16"), ("dcn_v2", "dcn_v2_v2_stacked"),
         ("dcn_v2", "dcn_v2_mix_parallel"), ("din", "din"), ("wide_deep", "wide_deep")]


@pytest.mark.parametrize("model,name", CASES)
def test_oracle_matches_golden(model, name):
    g = load_golden(name)
    pred, loss, grads = oracle_run(model, g)
    np.testing.assert_allclose(pred.numpy(), g["out"]["pred"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(loss.numpy(), g["out"]["loss"], rtol=0, atol=1e-13)
    assert set(grads) == set(g["grad"])
    for k, v in grads.items():
        np.testing.assert_allclose(v.numpy(), g["grad"][k], rtol=0, atol=1e-13, err_msg=k)


def test_golden_covers_edge_cases():
    g = load_golden("deepfm_d16")
    ids = g["in"]["ids"]
    assert (ids == 0).any() and (ids[1] == 0).all()            # padding id, all-padding sample
    assert ids[2, 5] == ids[2, 4] and len(set(ids[3:, 0])) == 1  # duplicates in/across samples
    # padding row gets no gradient (paddle.nn.Embedding(padding_idx=0))
    assert not g["grad"]["fm.embedding.weight"][0].any()
    assert not g["grad"]["fm.embedding_one.weight"][0].any()
    d = load_golden("din")
    assert d["in"]["lens"].min() == 1 and d["in"]["lens"].max() == d["in"]["hist_item"].shape[1]


def test_din_attention_grads_present():
    """the attention-unit linears are hidden from the reference's state_dict but do
    receive gradients; the golden file carries them under att.*"""
    g = load_golden("din")
    assert "att.linear_0.weight" in g["grad"] and np.abs(g["grad"]["att.linear_0.weight"]).max() > 0
