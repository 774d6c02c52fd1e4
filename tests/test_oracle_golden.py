"""Pins oracle/nets.py to the committed golden vectors (minted by executing the reference's own
net.py files on the paddle shim — tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from tests.util import load_golden, oracle_run

CASES = [("deepfm", "deepfm_d9"), ("deepfm", "deepfm_d16"), ("dcn_v2", "dcn_v2_v2_stacked"),
         ("dcn_v2", "dcn_v2_mix_parallel"), ("din", "din"), ("wide_deep", "wide_deep"), ("dlrm", "dlrm_pairs"), ("dlrm", "dlrm_self")]


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


def test_dlrm_self_interaction_diagonal_is_zero_not_the_self_dot():
    """dlrm/net.py:105-113: triu(Z,1) zeroes the diagonal BEFORE the MIN_FLOAT mask is added, so with
    self_interaction=True the N extra positions are selected but hold 0 — the golden (minted from the
    reference's code) pins that; a "correct" <T_i,T_i> would change the prediction."""
    import torch

    from oracle import nets

    T = torch.randn(3, 5, 4, dtype=torch.float64)
    R = nets.dot_interact(T, True)
    assert R.shape == (3, 4 + 15)
    iu = torch.triu_indices(5, 5, 0)
    diag = (iu[0] == iu[1])
    assert (R[:, 4:][:, diag] == 0).all() and (R[:, 4:][:, ~diag] != 0).all()
    assert torch.equal(R[:, :4], T[:, 4])
    g = load_golden("dlrm_self")
    stats = [k for k in g["param"] if k.endswith("._variance")]
    assert stats and all((g["param"][k] != 1).any() for k in stats)   # running stats moved
