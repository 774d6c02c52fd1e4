"""N>1 path on real GPUs (needs >= 2 visible devices; skipped otherwise): NCCL all-to-all + the CUDA
kernels must reproduce the single-process oracle on the global batch."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nets
from tests.test_sharded_cpu import D, Dn, F, FC, V, _free_port, _full_problem, _NoStep

B = 48      # divisible by 2, 4 and 8 ranks

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from paddlerec_b200 import functional as BF
        from paddlerec_b200 import sharded
        dev = torch.device("cuda", rank)
        p, ids, dense, label = _full_problem(B)
        torch.manual_seed(100 + rank)
        model = sharded.ShardedDeepFMLayer(V, D, Dn, F, FC, rank, world, device=dev)
        with torch.no_grad():
            sd = model.state_dict()
            for k, v in p.items():
                sd[k].copy_(v[rank::world] if k.startswith("fm.embedding") else v)
        per = B // world
        sl = slice(rank * per, (rank + 1) * per)
        pred = model(ids[sl].to(dev), dense[sl].to(dev))
        loss = BF.log_loss(pred, label[sl].to(dev)).mean()
        opt = sharded.DistributedOptimizer(_NoStep(), model, world)
        opt.scale_loss(loss).backward()
        opt.step()
        res = {"pred": pred.detach().cpu().numpy(),
               "dW": model.fm.table_grad_dense()[0].cpu().numpy(),
               "dW1": model.fm.table_grad_dense()[1].cpu().numpy()}
        for k, v in model.named_parameters():
            if v.grad is not None:
                res["g:" + k] = v.grad.cpu().numpy()
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_deepfm_nccl(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    p, ids, dense, label = _full_problem(B)
    pp = {k: v.double().requires_grad_(True) for k, v in p.items()}
    ids_ok = ids.clone()
    ids_ok[ids_ok >= V] = 0
    pred = nets.deepfm_forward(pp, [ids_ok[:, i:i + 1] for i in range(F)], dense.double(), len(FC))
    nets.log_loss(pred, label.double()).mean().backward()
    per = B // world
    for rank in range(world):
        r = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        np.testing.assert_allclose(r["pred"], pred.detach().numpy()[rank * per:(rank + 1) * per],
                                   rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(r["dW"], pp["fm.embedding.weight"].grad.numpy()[rank::world],
                                   rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(r["dW1"], pp["fm.embedding_one.weight"].grad.numpy()[rank::world],
                                   rtol=1e-4, atol=1e-7)
        for k in pp:
            if not k.startswith("fm.embedding"):
                np.testing.assert_allclose(r["g:" + k], pp[k].grad.numpy(), rtol=1e-4, atol=1e-7,
                                           err_msg=k)


def _dcn_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from paddlerec_b200 import functional as BF
        from paddlerec_b200 import sharded
        from paddlerec_b200.rank.dcn_v2 import net
        from tests.util import load_golden
        dev = torch.device("cuda", rank)
        g = load_golden("dcn_v2_v2_stacked")
        Vg, Dg = g["param"]["embedding.weight"].shape
        fc = [g["param"]["DNN_.linear_%d.weight" % i].shape[1] for i in range(2)]
        torch.manual_seed(3 + rank)
        model = net.DCN_V2Layer(Vg, Dg, 13, 26, fc, 2, True, False, 6, 4, device=dev)
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(torch.tensor(g["param"][k], dtype=torch.float32))
        sharded.shard_embeddings(model, rank, world)
        model.eval()
        ids = torch.tensor(g["in"]["ids"])
        Bg = ids.shape[0] // world * world
        per = Bg // world
        sl = slice(rank * per, (rank + 1) * per)
        dense = torch.tensor(g["in"]["dense"], dtype=torch.float32)
        label = torch.tensor(g["in"]["label"], dtype=torch.float32)
        pred = model(ids[sl].to(dev), dense[sl].to(dev))
        loss = BF.log_loss(pred, label[sl].to(dev)).sum() / Bg
        loss.backward()
        out = dict(pred=pred.detach().cpu().numpy(),
                   dW=model.embedding.grad_rows.to_dense().cpu().numpy())
        # one SGD step with ClipGradByGlobalNorm (models/rank/dcn_v2/dygraph_model.py:81-88): the
        # norm spans the dense gradients AND every rank's table shard (one scalar all-reduce)
        from paddlerec_b200 import optim
        inner = optim.SGD(0.5, model.parameters(), grad_clip=optim.ClipGradByGlobalNorm(DCN_CLIP))
        opt = sharded.DistributedOptimizer(inner, model, world)
        # loss above is already sum/Bg = the global-batch mean share of this rank
        opt.step()
        for k, v in model.state_dict().items():
            out["p:" + k] = v.detach().cpu().numpy()
        np.savez(os.path.join(out_dir, "dcn%d.npz" % rank), **out)
    finally:
        dist.destroy_process_group()


DCN_CLIP = 0.05


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_dcn_v2_nccl(world, tmp_path):
    """BASELINE config 3 in miniature: DCN-V2 with its table row-sharded over `world` GPUs, forward
    + gradients vs the oracle, then one clipped SGD step (global-norm clip across the shards)."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from tests.util import load_golden, slots, to_params
    mp.spawn(_dcn_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = load_golden("dcn_v2_v2_stacked")
    p = to_params(g["param"])
    ids = torch.tensor(g["in"]["ids"])
    Bg = ids.shape[0] // world * world
    dense = torch.tensor(g["in"]["dense"], dtype=torch.float64)[:Bg]
    label = torch.tensor(g["in"]["label"], dtype=torch.float64)[:Bg]
    pred = nets.dcn_v2_forward(p, slots(ids[:Bg]), dense, n_fc=2, cross_num=2, is_stacked=True,
                               use_low_rank_mixture=False)
    nets.log_loss(pred, label).mean().backward()
    per = Bg // world
    for rank in range(world):
        r = np.load(os.path.join(str(tmp_path), "dcn%d.npz" % rank))
        np.testing.assert_allclose(r["pred"], pred.detach().numpy()[rank * per:(rank + 1) * per],
                                   rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(r["dW"], p["embedding.weight"].grad.numpy()[rank::world],
                                   rtol=1e-4, atol=1e-7)
    norm = float(sum((v.grad ** 2).sum() for v in p.values() if v.grad is not None) ** 0.5)
    assert norm > 2 * DCN_CLIP
    scale = DCN_CLIP / max(norm, DCN_CLIP)
    r = [np.load(os.path.join(str(tmp_path), "dcn%d.npz" % k)) for k in range(world)]
    for k, v in p.items():
        if v.grad is None:
            continue
        want = (v.detach() - 0.5 * scale * v.grad).numpy()
        if k == "embedding.weight":
            for rank in range(world):
                np.testing.assert_allclose(r[rank]["p:" + k], want[rank::world], rtol=1e-4,
                                           atol=1e-6)
        else:
            for rank in range(1, world):   # replicas must stay bit-identical
                np.testing.assert_array_equal(r[0]["p:" + k], r[rank]["p:" + k])
            np.testing.assert_allclose(r[0]["p:" + k], want, rtol=1e-4, atol=1e-6, err_msg=k)
