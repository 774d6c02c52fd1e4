"""The plain-C restatement of the FM block (oracle/fm_ref.c) agrees with the torch oracle
(oracle/nets.py + autograd) — two independent CPU restatements of models/rank/deepfm/net.py:105-139."""
import ctypes
import os
import subprocess

import numpy as np
import torch

from oracle import nets

ORACLE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")


def _lib():
    subprocess.run(["make", "-s", "-C", ORACLE], check=True)
    return ctypes.CDLL(os.path.join(ORACLE, "_build", "libfm_ref.so"))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_c_restatement_matches_torch_oracle():
    lib = _lib()
    B, F, Dn, D, V = 37, 26, 13, 9, 120
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(0, V, (B, F), generator=g)
    ids[1] = 0
    dense = torch.rand(B, Dn, generator=g)
    p = {"fm.embedding.weight": (torch.randn(V, D, generator=g) * 0.1).requires_grad_(True),
         "fm.embedding_one.weight": (torch.randn(V, 1, generator=g) * 0.1).requires_grad_(True),
         "fm.dense_w": (torch.randn(1, Dn, D, generator=g) * 0.1).requires_grad_(True),
         "fm.dense_w_one": (torch.randn(Dn, generator=g) * 0.1).requires_grad_(True)}
    y1, y2, feat = nets.deepfm_fm(p, [ids[:, i:i + 1] for i in range(F)], dense)
    A = torch.randn(B, F + Dn, D, generator=g)
    g1 = torch.randn(B, generator=g)
    g2 = torch.randn(B, generator=g)
    ((feat * A).sum() + (y1.reshape(-1) * g1).sum() + (y2.reshape(-1) * g2).sum()).backward()

    W = p["fm.embedding.weight"].detach().numpy().copy()
    W1 = p["fm.embedding_one.weight"].detach().numpy().reshape(-1).copy()
    dw = p["fm.dense_w"].detach().numpy().reshape(Dn, D).copy()
    dw1 = p["fm.dense_w_one"].detach().numpy().copy()
    ids_np, dense_np = ids.numpy().copy(), dense.numpy().copy()
    cfeat = np.zeros((B, F + Dn, D), np.float32)
    cy1, cy2 = np.zeros(B, np.float32), np.zeros(B, np.float32)
    lib.fm_ref_fwd(_p(W), _p(W1), _p(ids_np), _p(dense_np), _p(dw), _p(dw1), _p(cfeat), _p(cy1),
                   _p(cy2), ctypes.c_int64(B), F, Dn, D, ctypes.c_int64(0))
    np.testing.assert_allclose(cfeat, feat.detach().numpy(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(cy1, y1.detach().numpy().reshape(-1), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(cy2, y2.detach().numpy().reshape(-1), rtol=2e-5, atol=2e-6)
    dW, dW1 = np.zeros((V, D)), np.zeros(V)
    ddw, ddw1 = np.zeros((Dn, D)), np.zeros(Dn)
    An, g1n, g2n = A.numpy().copy(), g1.numpy().copy(), g2.numpy().copy()
    lib.fm_ref_bwd(_p(ids_np), _p(dense_np), _p(cfeat), _p(An), _p(g1n), _p(g2n), _p(dW), _p(dW1),
                   _p(ddw), _p(ddw1), ctypes.c_int64(B), F, Dn, D, ctypes.c_int64(0))
    np.testing.assert_allclose(dW, p["fm.embedding.weight"].grad.numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(dW1, p["fm.embedding_one.weight"].grad.numpy().reshape(-1), rtol=2e-5,
                               atol=2e-6)
    np.testing.assert_allclose(ddw, p["fm.dense_w"].grad.numpy()[0], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ddw1, p["fm.dense_w_one"].grad.numpy(), rtol=2e-5, atol=2e-6)
    assert not dW[0].any()


def _interact_lib():
    subprocess.run(["make", "-s", "-C", ORACLE], check=True)
    return ctypes.CDLL(os.path.join(ORACLE, "_build", "libinteract_ref.so"))


def test_c_dot_interact_matches_torch_oracle():
    """oracle/interact_ref.c (loops over the triangle, double accumulation) against
    oracle/nets.dot_interact (bmm + triu indices + autograd) — dlrm/net.py:97-115 restated twice."""
    lib = _interact_lib()
    for (B, N, d, self_i) in [(4, 27, 16, 0), (3, 27, 8, 1), (2, 2, 1, 0), (5, 6, 3, 1)]:
        g = torch.Generator().manual_seed(B * N + d)
        T = torch.randn(B, N, d, generator=g)
        Tq = T.double().requires_grad_(True)
        R = nets.dot_interact(Tq, bool(self_i))
        dR = torch.randn(R.shape, generator=g, dtype=torch.float64)
        R.backward(dR)
        Tn = T.numpy().copy()
        cR = np.zeros(R.shape, np.float64)
        cdT = np.zeros((B, N, d), np.float64)
        lib.dot_interact_ref_fwd(_p(Tn), _p(cR), ctypes.c_int64(B), N, d, self_i)
        dRn = dR.numpy().copy()
        lib.dot_interact_ref_bwd(_p(Tn), _p(dRn), _p(cdT), ctypes.c_int64(B), N, d, self_i)
        np.testing.assert_allclose(cR, R.detach().numpy(), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(cdT, Tq.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_c_hash_keys_matches_numpy_oracle():
    from oracle import readers
    lib = _interact_lib()
    rng = np.random.default_rng(3)
    n = 5000
    keys = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, n, dtype=np.uint64)
    keys[:3] = np.array([0, 1, (1 << 64) - 1], np.uint64)
    slots = rng.integers(0, 26, n).astype(np.int32)
    for V, rz in [(1000001, 1), (2, 1), (97, 0), ((1 << 40) + 7, 1)]:
        for sl in (None, slots):
            rows = np.zeros(n, np.int64)
            lib.hash_keys_ref(_p(keys), _p(sl) if sl is not None else None, ctypes.c_int64(n),
                              ctypes.c_uint64(V), rz, _p(rows))
            assert np.array_equal(rows, readers.hash_keys(keys, V, sl, bool(rz)))
    # splitmix64's published first output for state 0: mix(0 + golden) = 0xE220A8397B1DCDAF
    one = np.zeros(1, np.int64)
    k = np.array([0x9E3779B97F4A7C15], np.uint64)
    lib.hash_keys_ref(_p(k), None, ctypes.c_int64(1), ctypes.c_uint64(1 << 63), 0, _p(one))
    assert int(one[0]) == 0xE220A8397B1DCDAF % (1 << 63)
