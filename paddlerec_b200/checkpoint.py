"""Checkpoints in the reference's on-disk layout (tools/utils/save_load.py:25-46):
`<model_save_path>/<epoch>/rec.pdparams` and `rec.pdopt`.

`.pdparams` is written the way `paddle.save(layer.state_dict(), path)` writes a state_dict in
Paddle 2.x (public behaviour, restated — verify on a machine that has Paddle): ONE pickle
(protocol 4) of a plain dict `structured name -> numpy.ndarray`, plus the entry
`"StructuredToParameterName@@"` mapping every structured name to the parameter's internal name
(we have no internal names, so the structured name stands in).  `paddle.load` accepts exactly this
and `Layer.set_state_dict` matches by structured name, so a Paddle process can read our file and we
can read Paddle's.  Names and layouts are the reference's (`Linear.weight [in,out]`,
`fm.embedding.weight`, BatchNorm `_mean` / `_variance`): nn.FusedTable exports its slots as the
reference's two tables.

`.pdopt` holds what a resume needs — step count, LR-scheduler epoch, dense Adam/SGD state and the
row-wise moments of the sparse tables — keyed by structured names.  Paddle's own `.pdopt` keys
accumulators by internal parameter names, which do not exist here, so this file is ours.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict

import numpy as np
import torch

NAME_TABLE_KEY = "StructuredToParameterName@@"
_TORCH_ZIP_MAGIC = b"PK\x03\x04"


def to_numpy_state(state_dict) -> Dict[str, np.ndarray]:
    out = {}
    for k, v in state_dict.items():
        if k == NAME_TABLE_KEY:
            continue
        out[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
    return out


def save_pdparams(state_dict, path: str) -> None:
    saved = to_numpy_state(state_dict)
    saved[NAME_TABLE_KEY] = {k: k for k in saved}
    tmp = path + ".tmp.%d" % os.getpid()
    with open(tmp, "wb") as fh:
        pickle.dump(saved, fh, protocol=4)
    os.replace(tmp, path)


def load_pdparams(path: str) -> Dict[str, np.ndarray]:
    """-> structured name -> ndarray.  Reads the pickle layout above (ours or Paddle's); a legacy
    torch-zip checkpoint of earlier revisions of this repo is still accepted."""
    with open(path, "rb") as fh:
        head = fh.read(4)
    if head == _TORCH_ZIP_MAGIC:
        return to_numpy_state(torch.load(path, map_location="cpu"))
    with open(path, "rb") as fh:
        obj = pickle.load(fh)
    if not isinstance(obj, dict):
        raise ValueError("%s does not hold a state_dict" % path)
    obj.pop(NAME_TABLE_KEY, None)
    out = {}
    for k, v in obj.items():
        if isinstance(v, tuple) and len(v) == 2 and isinstance(v[1], np.ndarray):
            v = v[1]              # (name, ndarray) pairs of paddle's tensor reducer
        out[k] = np.asarray(v)
    return out


def set_state_dict(net: torch.nn.Module, state: Dict[str, np.ndarray], strict: bool = True):
    """Layer.set_state_dict: match by structured name, check shapes, copy on the module's device.

    A module may declare `optional_state_prefixes`: keys of ITS state_dict that a reference-produced
    checkpoint does not contain.  DIN's attention-unit linears are the case (the
    reference's sub-layer name collision drops them from state_dict(), so a Paddle rec.pdparams
    has no `attention.linear_*`; they keep their seeded initialisation, as in the reference)."""
    dev = next((p.device for p in net.parameters()), torch.device("cpu"))
    tensors = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in state.items()}
    optional = tuple(getattr(net, "optional_state_prefixes", ()))
    if strict and optional:
        own = net.state_dict()
        for k, v in own.items():
            if k not in tensors and k.startswith(optional):
                tensors[k] = v
    return net.load_state_dict(tensors, strict=strict)


# ---- optimizer ------------------------------------------------------------------------------------
def _names(net: torch.nn.Module) -> Dict[int, str]:
    return {id(p): n for n, p in net.named_parameters()}


def optimizer_state(optimizer, net: torch.nn.Module) -> dict:
    names = _names(net)
    out = {"class": type(optimizer).__name__, "step_count": int(optimizer.step_count), "dense": {},
           "sparse": {}}
    sched = getattr(optimizer, "_lr", None)
    if hasattr(sched, "last_epoch"):
        out["lr_last_epoch"] = int(sched.last_epoch)
    topt = getattr(optimizer, "_torch", None)
    if topt is not None:
        for p, st in topt.state.items():
            out["dense"][names[id(p)]] = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v)
                                          for k, v in st.items()}
    for p in optimizer._sparse:
        name = names[id(p)]
        if hasattr(optimizer, "moments") and (getattr(p, "inslot_moments", None) is not None
                                              or id(p) in getattr(optimizer, "_m", {})):
            m, v = optimizer.moments(p)
            out["sparse"][name] = {"moment1": m.detach().cpu().numpy(), "moment2": v.detach().cpu().numpy()}
        elif hasattr(optimizer, "_g2") and id(p) in optimizer._g2:
            out["sparse"][name] = {"g2sum": optimizer._g2[id(p)].detach().cpu().numpy()}
    return out


def load_optimizer_state(optimizer, net: torch.nn.Module, state: dict) -> None:
    if state.get("class") != type(optimizer).__name__:
        raise ValueError("checkpoint holds %s state, optimizer is %s"
                         % (state.get("class"), type(optimizer).__name__))
    by_name = {n: p for n, p in net.named_parameters()}
    optimizer.step_count = int(state["step_count"])
    sched = getattr(optimizer, "_lr", None)
    if "lr_last_epoch" in state and hasattr(sched, "last_epoch"):
        sched.last_epoch = int(state["lr_last_epoch"])
    topt = getattr(optimizer, "_torch", None)
    for name, st in state["dense"].items():
        p = by_name[name]
        topt.state[p] = {k: (torch.as_tensor(v).to(p.device) if isinstance(v, np.ndarray) else v)
                         for k, v in st.items()}
    for name, st in state["sparse"].items():
        p = by_name[name]
        if "moment1" in st:
            m, v = optimizer.moments(p)
            m.copy_(torch.from_numpy(st["moment1"]).to(p.device))
            v.copy_(torch.from_numpy(st["moment2"]).to(p.device))
        elif "g2sum" in st:
            optimizer.g2sum(p).copy_(torch.from_numpy(st["g2sum"]).to(p.device))


def save_pdopt(optimizer, net, path: str) -> None:
    tmp = path + ".tmp.%d" % os.getpid()
    with open(tmp, "wb") as fh:
        pickle.dump(optimizer_state(optimizer, net), fh, protocol=4)
    os.replace(tmp, path)


def load_pdopt(optimizer, net, path: str) -> None:
    with open(path, "rb") as fh:
        load_optimizer_state(optimizer, net, pickle.load(fh))
