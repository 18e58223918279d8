"""Prompt path of a request: feature dict -> embedding ids -> [B, 128, F] conditioning.

Mirrors the reference's two pieces with the same names, argument meaning and error behaviour:
  * ``feature_dict_to_embedding_ids(feature_dict, feature_yaml)``  -- mug/util.py:62-84 (host logic, called per request by
    webui.py:186-193 / scripts/mapping.py:44-52); ``count_beatmap_features`` -- mug/util.py:50-60,86-90;
  * ``PromptEmbedder`` -- mug/cond/feature.py:8-21 (``BeatmapFeatureEmbedder``): nn.Embedding lookup + "b f h -> b h f",
    here one gather kernel of libmugd (MUGD_OP_EMBED) on the weights' device.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Sequence

import torch

from . import lib as L_


def count_beatmap_features_embedding(x: Dict[str, Any]) -> int:
    """rows one feature occupies in the table: slot 0 = "missing", then the value bins (mug/util.py:50-60)"""
    kind = x["type"]
    if kind == "numeric":
        return int(math.ceil((x["max"] - x["min"]) / x["interval"])) + 1
    if kind == "category":
        return len(x["category"]) + 1
    if kind == "bool":
        return 3
    raise ValueError(str(x))


def count_beatmap_features(feature_yaml: Sequence[Dict[str, Any]]) -> int:
    """table height: every feature contributes its rows once per ``count`` slot (mug/util.py:86-90)"""
    return sum(count_beatmap_features_embedding(x) * x.get("count", 1) for x in feature_yaml)


def feature_dict_to_embedding_ids(feature_dict: Dict[str, Any], feature_yaml: Sequence[Dict[str, Any]]) -> List[int]:
    """One table row id per feature slot, in yaml order (mug/util.py:62-84).

    Numeric values are clamped to [min, max] and binned by ``interval`` (truncation); bools index by their int value; a
    category is its position in the list -- an unknown category raises ValueError exactly like the reference's ``list.index``.
    Bin 0 of every feature means "not given".  A feature with ``count`` n fills n consecutive slots with the same bin, each
    slot owning its own block of rows.
    """
    ids: List[int] = []
    base = 0
    for x in feature_yaml:
        value = feature_dict.get(x["name"], None)
        if value is None:
            bin_ = 0
        else:
            if x["type"] == "numeric":
                value = max(x["min"], min(x["max"], value))
                bin_ = int((value - x["min"]) / x["interval"])
            elif x["type"] == "bool":
                bin_ = value
            else:
                bin_ = x["category"].index(value)
            bin_ += 1
        width = count_beatmap_features_embedding(x)
        for _ in range(x.get("count", 1)):
            ids.append(bin_ + base)
            base += width
    return ids


class PromptEmbedder:
    """``model.model.cond_stage_model``: ids [B, F] (any real dtype, truncated like ``x.long()``) -> [B, H, F] fp32."""

    def __init__(self, engine, weight: torch.Tensor):
        self.engine = engine
        self.weight = weight.detach().to(device=engine.device, dtype=torch.float32).contiguous()
        self.n_embed, self.embed_dim = self.weight.shape

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 2:
            raise ValueError(f"expected ids of shape [B, F], got {tuple(x.shape)}")
        ids = x.long()
        lo, hi = int(ids.min()), int(ids.max())
        if lo < 0 or hi >= self.n_embed:
            raise IndexError("index out of range in self")          # torch.nn.functional.embedding's message
        B, F = ids.shape
        ids32 = ids.to(device=self.engine.device, dtype=torch.int32).contiguous()
        out = torch.empty(B, self.embed_dim, F, device=self.engine.device, dtype=torch.float32)
        e = L_.Embed()
        e.table, e.ids, e.out = self.weight.data_ptr(), ids32.data_ptr(), out.data_ptr()
        e.B, e.F, e.H, e.n_embed = B, F, self.embed_dim, self.n_embed
        op = L_.make_op(L_.OP_EMBED, e)
        with self.engine.lock:
            import ctypes as C
            L_.check(self.engine.lib.mugd_op_run(self.engine.handle, C.byref(op), torch.cuda.current_stream().cuda_stream), "embed")
        return out

    forward = __call__
