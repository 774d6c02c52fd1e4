"""Config loading and the dygraph training loop — the caller side of the hot path, shaped like the
reference's tools/trainer.py:49-223 and tools/utils/utils_single.py (load_yaml :131-136,
get_all_inters_from_yaml :57-86, create_data_loader :89-113, load_dy_model_class :116-120) so that
`python tools/trainer.py -m <config.yaml> [-o key=value ...]` behaves the same way.
"""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import logging
import os
import sys
import time
from typing import Dict, List

import numpy as np
import torch
import yaml

from . import checkpoint

logging.basicConfig(format="%(asctime)s - %(levelname)s - %(message)s", level=logging.INFO)
logger = logging.getLogger("paddlerec_b200")


# ---- yaml -> flat dotted dict (utils_single.py:57-86,131-136) -----------------------------------
def flatten_yaml(envs: dict, filters=("workspace", "runner", "hyper_parameters")) -> Dict[str, object]:
    flat: Dict[str, object] = {}

    def walk(prefix: List[str], node: dict):
        for k, v in node.items():
            if isinstance(v, dict):
                walk(prefix + [k], v)
            elif k in ("dataset", "phase", "runner") and isinstance(v, list):
                for item in v:
                    if item.get("name") is None:
                        raise ValueError("name must be in dataset list. ", v)
                    walk(prefix + [k, item["name"]], item)
            else:
                flat[".".join(prefix + [k])] = v

    walk([], envs)
    return {k: v for k, v in flat.items() if any(k.startswith(f) for f in filters)}


def load_yaml(yaml_file: str) -> Dict[str, object]:
    with open(yaml_file, "r") as fh:
        envs = yaml.safe_load(fh)
    return flatten_yaml(envs)


def apply_overrides(config: dict, opts) -> None:
    """`-o key=value`: typed by the existing value (trainer.py:55-65)."""
    for parameter in opts or []:
        parameter = parameter.strip()
        key, _, value = parameter.partition("=")
        if isinstance(config.get(key), bool):
            config[key] = value in ("True", "true", "1")
        elif isinstance(config.get(key), int):
            config[key] = int(value)
        elif isinstance(config.get(key), float):
            config[key] = float(value)
        elif isinstance(config.get(key), list):
            config[key] = yaml.safe_load(value)
        else:
            config[key] = value


# ---- plugin loading (utils_single.py:116-120) ---------------------------------------------------
def _import_from_dir(abs_dir: str, module: str):
    """Import <abs_dir>/<module>.py.  Our own model dirs are packages (relative imports); a foreign
    plugin dir is imported the reference's way (sys.path + top-level name)."""
    pkg_root = os.path.dirname(os.path.abspath(__file__))
    abs_dir = os.path.abspath(abs_dir)
    if abs_dir.startswith(pkg_root + os.sep):
        rel = os.path.relpath(abs_dir, os.path.dirname(pkg_root)).replace(os.sep, ".")
        return importlib.import_module(rel + "." + module)
    if abs_dir not in sys.path:
        sys.path.append(abs_dir)
    return importlib.import_module(module)


def load_dy_model_class(abs_dir: str):
    return _import_from_dir(abs_dir, "dygraph_model").DygraphModel()


def _collate(samples):
    """Stack each slot: list of per-sample arrays -> list of [B, ...] tensors."""
    return [torch.from_numpy(np.stack([s[i] for s in samples])) for i in range(len(samples[0]))]


def create_data_loader(config, mode="train", rank=0, world_size=1):
    if mode == "train":
        data_dir = config.get("runner.train_data_dir")
        batch_size = config.get("runner.train_batch_size")
        reader_path = config.get("runner.train_reader_path", "reader")
    else:
        data_dir = config.get("runner.test_data_dir")
        batch_size = config.get("runner.infer_batch_size")
        reader_path = config.get("runner.infer_reader_path", "reader")
    abs_dir = config["config_abs_dir"]
    data_dir = os.path.join(abs_dir, data_dir)
    file_list = [os.path.join(data_dir, x) for x in sorted(os.listdir(data_dir))]
    reader_type = config.get("runner.reader_type", "DataLoader")
    if reader_type == "PackedReader":
        # native parser (dataio.py / libb200rec_io.so): Criteo `slot:value` files -> packed
        # (label, ids, dense) batches in pinned memory; same sample order and drop-last rule
        from . import dataio

        fmt = config.get("runner.packed_format", "slot_text")
        if fmt == "din":     # behaviour logs: length-grouped, per-batch padding (dinReader.py)
            return dataio.DinBatchReader(file_list, batch_size=batch_size,
                                         threads=int(config.get("runner.reader_threads", 0)),
                                         pin_memory=torch.cuda.is_available())
        schemas = {"criteo": dataio.CRITEO, "criteo_dcn_v2": dataio.CRITEO_DCN_V2}
        schema_name = config.get("runner.packed_schema", "criteo")
        if schema_name not in schemas:
            raise ValueError("runner.packed_schema %r is not one of %s" % (schema_name, sorted(schemas)))
        return dataio.PackedBatchReader(
            file_list, batch_size=batch_size, fmt=fmt, schema=schemas[schema_name], drop_last=True,
            threads=int(config.get("runner.reader_threads", 0)),
            pin_memory=torch.cuda.is_available(), rank=rank, world_size=world_size,
            shard_files=bool(config.get("runner.use_fleet", False)))
    if reader_type != "DataLoader":
        raise ValueError("runner.reader_type %r is not supported (DataLoader | PackedReader)" % reader_type)
    reader = _import_from_dir(abs_dir, reader_path)
    dataset = reader.RecDataset(file_list, config=config)
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, drop_last=True,
                                       collate_fn=_collate,
                                       num_workers=int(config.get("runner.num_workers", 0)))


# ---- checkpoint (save_load.py:25-46) ------------------------------------------------------------
def save_model(net, optimizer, model_path, epoch_id, prefix="rec"):
    """<model_path>/<epoch>/rec.pdparams (+ rec.pdopt) — the layout of save_load.py:25-31, the
    .pdparams in paddle.save's state_dict pickle format (checkpoint.py)."""
    path = os.path.join(model_path, str(epoch_id))
    os.makedirs(path, exist_ok=True)
    checkpoint.save_pdparams(net.state_dict(), os.path.join(path, prefix + ".pdparams"))
    if optimizer is not None:
        checkpoint.save_pdopt(optimizer, net, os.path.join(path, prefix + ".pdopt"))
    logger.info("Already save model in %s", path)


def load_model(model_path, net, prefix="rec", optimizer=None):
    """save_load.py:41-46; with `optimizer`, also restores rec.pdopt when it is there (resume)."""
    logger.info("start load model from %s", model_path)
    checkpoint.set_state_dict(net, checkpoint.load_pdparams(os.path.join(model_path, prefix + ".pdparams")))
    opt_path = os.path.join(model_path, prefix + ".pdopt")
    if optimizer is not None and os.path.exists(opt_path):
        checkpoint.load_pdopt(optimizer, net, opt_path)


class DevicePrefetcher:
    """Iterates a host-batch iterable and yields the batches as DEVICE tensors (the output of
    `dy_model_class.create_feeds`), uploading batch i+1 on a copy stream while the caller's step i
    runs on the compute stream — pinned host memory makes the copy asynchronous, so the copy engine
    works beside the kernels instead of in front of them.  `train_forward` accepts the yielded
    tuple unchanged (create_feeds is a no-op on device tensors).  `peek()` returns the batch that
    the next iteration will yield (the sharded model plans its id exchange from it)."""

    def __init__(self, batches, dy_model_class, config):
        self._it = iter(batches)
        self._dm, self._config = dy_model_class, config
        self._copy = torch.cuda.Stream()
        self._next = None
        self._upload()

    def _upload(self):
        try:
            host = next(self._it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self._copy):
            self._next = tuple(self._dm.create_feeds(host, self._config))

    def peek(self):
        return self._next

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        cur = torch.cuda.current_stream()
        cur.wait_stream(self._copy)
        batch = self._next
        for t in batch:
            if isinstance(t, torch.Tensor):
                t.record_stream(cur)
        self._upload()
        return batch


# ---- the loop (trainer.py:49-223) ---------------------------------------------------------------
def parse_args(argv=None):
    p = argparse.ArgumentParser("PaddleRec-shaped dygraph trainer on the b200rec engine")
    p.add_argument("-m", "--config_yaml", type=str, required=True)
    p.add_argument("-o", "--opt", nargs="*", type=str)
    args = p.parse_args(argv)
    args.abs_dir = os.path.dirname(os.path.abspath(args.config_yaml))
    args.config_yaml = os.path.abspath(args.config_yaml)
    return args


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("paddlerec_b200 runs on a CUDA device only (no CPU fallback)")


def train(config: dict, max_steps=None, save=True):
    """Returns a list of per-step python floats (loss) and the final metric values."""
    _require_cuda()
    dy_model_class = load_dy_model_class(config["config_abs_dir"])
    seed = int(config.get("runner.seed", 12345))
    torch.manual_seed(seed)
    epochs = int(config.get("runner.epochs"))
    print_interval = int(config.get("runner.print_interval", 10))
    model_save_path = config.get("runner.model_save_path", "model_output")
    model_init_path = config.get("runner.model_init_path", None)

    dy_model = dy_model_class.create_model(config)
    optimizer = dy_model_class.create_optimizer(dy_model, config)
    from . import tower
    tower.set_overlap_dw(True)          # the optimizers wait for the side-stream dW GEMMs
    if model_init_path is not None:     # warm start (trainer.py:106-107); also resumes rec.pdopt
        load_model(model_init_path, dy_model, optimizer=optimizer)
    train_dataloader = create_data_loader(config, "train")

    losses = []
    metric_values = {}
    step_num = 0
    for epoch_id in range(int(config.get("last_epoch", -1)) + 1, epochs):
        dy_model.train()
        metric_list, metric_list_name = dy_model_class.create_metrics()
        train_reader_cost = train_run_cost = 0.0
        total_samples = 0
        reader_start = time.time()
        for batch_id, batch in enumerate(train_dataloader):
            train_reader_cost += time.time() - reader_start
            optimizer.clear_grad()
            train_start = time.time()
            batch_size = len(batch[0])
            loss, metric_list, tensor_print_dict = dy_model_class.train_forward(
                dy_model, metric_list, batch, config)
            loss.backward()
            optimizer.step()
            total_samples += batch_size
            if batch_id % print_interval == 0:
                # the only host sync of the loop
                metric_str = "".join("%s:%.6f, " % (n, m.accumulate())
                                     for n, m in zip(metric_list_name, metric_list))
                tensor_str = "".join("%s:%s," % (k, str(float(v.detach() if hasattr(v, "detach") else v))) for k, v in
                                     (tensor_print_dict or {}).items())
                train_run_cost += time.time() - train_start
                logger.info(
                    "epoch: %d, batch_id: %d, %s%s avg_reader_cost: %.5f sec, avg_batch_cost: %.5f "
                    "sec, avg_samples: %.5f, ips: %.5f ins/s", epoch_id, batch_id, metric_str,
                    tensor_str, train_reader_cost / print_interval,
                    (train_reader_cost + train_run_cost) / print_interval,
                    total_samples / print_interval,
                    total_samples / (train_reader_cost + train_run_cost + 0.0001))
                train_reader_cost = train_run_cost = 0.0
                total_samples = 0
            else:
                train_run_cost += time.time() - train_start
            losses.append(loss.detach())
            reader_start = time.time()
            step_num += 1
            if max_steps is not None and step_num >= max_steps:
                break
        metric_values = {n: m.accumulate() for n, m in zip(metric_list_name, metric_list)}
        logger.info("epoch: %d done, %s", epoch_id,
                    ", ".join("%s: %.6f" % kv for kv in metric_values.items()))
        if save:
            save_model(dy_model, optimizer, os.path.join(config["config_abs_dir"], model_save_path)
                       if not os.path.isabs(model_save_path) else model_save_path, epoch_id)
        if max_steps is not None and step_num >= max_steps:
            break
    return [float(l) for l in losses], metric_values, dy_model


def infer(config: dict, max_steps=None):
    """tools/infer.py:48-197: for every saved epoch in [infer_start_epoch, infer_end_epoch) load
    <infer_load_path>/<epoch>/rec.pdparams, run infer_forward over the test data in eval mode and
    report the metrics; the metric objects are created once and reset after every epoch only when
    `runner.use_auc` is set (infer.py:180-181).  Returns {epoch: {metric: value}}."""
    _require_cuda()
    dy_model_class = load_dy_model_class(config["config_abs_dir"])
    torch.manual_seed(12345)
    print_interval = int(config.get("runner.print_interval", 10))
    load_path = config.get("runner.infer_load_path", "model_output")
    if not os.path.isabs(load_path):
        load_path = os.path.join(config["config_abs_dir"], load_path)
    start_epoch = int(config.get("runner.infer_start_epoch", 0))
    end_epoch = int(config.get("runner.infer_end_epoch", 10))
    dy_model = dy_model_class.create_model(config)
    test_dataloader = create_data_loader(config, "test")
    metric_list, metric_list_name = dy_model_class.create_metrics()
    results, step_num = {}, 0
    for epoch_id in range(start_epoch, end_epoch):
        logger.info("load model epoch %d", epoch_id)
        load_model(os.path.join(load_path, str(epoch_id)), dy_model)
        dy_model.eval()
        seen = 0
        interval_begin = time.time()
        for batch_id, batch in enumerate(test_dataloader):
            batch_size = len(batch[0])
            metric_list, tensor_print_dict = dy_model_class.infer_forward(dy_model, metric_list, batch,
                                                                          config)
            seen += batch_size
            if batch_id % print_interval == 0:
                metric_str = "".join("%s: %.6f," % (n, m.accumulate())
                                     for n, m in zip(metric_list_name, metric_list))
                tensor_str = "".join("%s:%s," % (k, str(float(v.detach() if hasattr(v, "detach") else v))) for k, v in
                                     (tensor_print_dict or {}).items())
                logger.info("epoch: %d, batch_id: %d, %s%s ips: %.2f ins/s", epoch_id, batch_id,
                            metric_str, tensor_str,
                            print_interval * batch_size / (time.time() + 0.0001 - interval_begin))
                interval_begin = time.time()
            step_num += 1
            if max_steps is not None and step_num >= max_steps:
                break
        if seen == 0:
            raise RuntimeError("test_dataloader is null, please ensure batch size < dataset size!")
        results[epoch_id] = {n: m.accumulate() for n, m in zip(metric_list_name, metric_list)}
        logger.info("epoch: %d done, %s", epoch_id,
                    ", ".join("%s: %.6f" % kv for kv in results[epoch_id].items()))
        if config.get("runner.use_auc", False):
            for m in metric_list:
                m.reset()
        if max_steps is not None and step_num >= max_steps:
            break
    return results


def main(argv=None, mode="train"):
    args = parse_args(argv)
    config = load_yaml(args.config_yaml)
    config["yaml_path"] = args.config_yaml
    config["config_abs_dir"] = args.abs_dir
    apply_overrides(config, args.opt)
    logger.info("**************common.configs**********")
    for k in ("runner.use_gpu", "runner.train_batch_size", "runner.epochs", "runner.print_interval"):
        logger.info("%s: %s", k, config.get(k))
    return train(config) if mode == "train" else infer(config)


if __name__ == "__main__":
    main()
