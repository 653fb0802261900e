"""Dict contracts of the path, as ONE schema table: for every dict the reference passes around (its lib/type_hinting.py
names them LogitData, CategoricalData, AggData, MatchedData) the keys with their layouts.  The ``TypedDict`` classes the
reference's annotations use are generated from the table, and ``validate`` checks a dict against it (symbolic dims must
agree across keys); the tests run it on the oracle's and the kernels' results."""
from __future__ import annotations

import typing
from typing import Dict, Mapping, Tuple

import torch

# name -> key -> (symbolic shape, dtype).  b frames, C classes incl. background, K = C-1, N instances,
# M matched pairs, S = up-sampling factor of the heads (low-resolution variant).
SCHEMAS: Dict[str, Dict[str, Tuple[Tuple[str, ...], torch.dtype]]] = {
    "LogitData": {                                    # network heads, full resolution (lib/pose_regressor.py:706-741)
        "mask": (("b", "C", "h", "w"), torch.float32),
        "quaternion": (("b", "4K", "h", "w"), torch.float32),
        "scales": (("b", "3K", "h", "w"), torch.float32),
        "z": (("b", "K", "h", "w"), torch.float32),
        "xy": (("b", "2K", "h", "w"), torch.float32),
    },
    "LowResLogitData": {                              # the heads' 1x1-conv outputs before the x S up-sampling
        "mask": (("b", "C", "h/S", "w/S"), torch.float32),
        "quaternion": (("b", "4K", "h/S", "w/S"), torch.float32),
        "scales": (("b", "3K", "h/S", "w/S"), torch.float32),
        "z": (("b", "K", "h/S", "w/S"), torch.float32),
        "xy": (("b", "2K", "h/S", "w/S"), torch.float32),
    },
    "CategoricalData": {                              # after class compression (lib/gpu_tensor_funcs.py:52-99)
        "mask": (("b", "h", "w"), torch.int64),
        "quaternion": (("b", "4", "h", "w"), torch.float32),
        "scales": (("b", "3", "h", "w"), torch.float32),
        "z": (("b", "h", "w"), torch.float32),
        "xy": (("b", "2", "h", "w"), torch.float32),
    },
    "AggData": {                                      # per instance (lib/aggregation_layer.py:61-158, hough voting, RT)
        "class_ids": (("N",), torch.int64),
        "sample_ids": (("N",), torch.int64),
        "instance_masks": (("N", "h", "w"), torch.float32),
        "quaternion": (("N", "4"), torch.float32),
        "scales": (("N", "3"), torch.float32),
        "z": (("N", "1"), torch.float32),
        "xy": (("N", "2"), torch.float32),            # [N,2,h,w] between aggregation and voting
        "R": (("N", "3", "3"), torch.float32),
        "T": (("N", "3"), torch.float32),
        "RT": (("N", "4", "4"), torch.float32),
    },
    "MatchedData": {                                  # ground truth stacked on its matched prediction (lib/matching.py:226-325)
        "class_ids": (("M",), torch.int64),
        "sample_ids": (("M",), torch.int64),
        "symmetric_ids": (("M",), torch.int64),
        "instance_masks": (("2", "M", "h", "w"), torch.float32),
        "quaternion": (("2", "M", "4"), torch.float32),
        "scales": (("2", "M", "3"), torch.float32),
        "z": (("2", "M", "1"), torch.float32),
        "xy": (("2", "M", "2"), torch.float32),
        "R": (("2", "M", "3", "3"), torch.float32),
        "T": (("2", "M", "3"), torch.float32),
        "RT": (("2", "M", "4", "4"), torch.float32),
    },
}


def _typed(name: str):
    return typing.TypedDict(name, {k: torch.Tensor for k in SCHEMAS[name]}, total=False)


LogitData = _typed("LogitData")
LowResLogitData = _typed("LowResLogitData")
CategoricalData = _typed("CategoricalData")
AggData = _typed("AggData")
MatchedData = _typed("MatchedData")


def validate(data: Mapping[str, torch.Tensor], schema: str, **dims: int) -> Dict[str, int]:
    """Checks every key of ``data`` that ``schema`` knows: rank, dtype, literal dims, and that symbolic dims (``b``, ``h``,
    ``N`` ...; ``4K`` = 4*K, ``h/S`` = h//S) agree across keys and with ``dims``.  Returns the resolved dims."""
    known = dict(dims)
    for key, (shape, dtype) in SCHEMAS[schema].items():
        if key not in data:
            continue
        t = data[key]
        if key == "xy" and schema == "AggData" and t.dim() == 4:
            shape = ("N", "2", "h", "w")
        if t.dim() != len(shape):
            raise ValueError(f"{schema}['{key}'] must have {len(shape)} dims {shape}, got {tuple(t.shape)}")
        if t.dtype != dtype:
            raise ValueError(f"{schema}['{key}'] must be {dtype}, got {t.dtype}")
        for sym, size in zip(shape, t.shape):
            if sym.isdigit():
                want = int(sym)
            elif sym[0].isdigit():                        # e.g. 4K
                if "K" not in known:
                    known["K"] = size // int(sym[0])
                want = int(sym[0]) * known["K"]
            elif "/" in sym:                              # e.g. h/S
                base, div = sym.split("/")
                if base in known and div in known:
                    want = known[base] // known[div]
                else:
                    known.setdefault(sym, size)
                    want = known[sym]
            else:
                known.setdefault(sym, size)
                want = known[sym]
            if size != want:
                raise ValueError(f"{schema}['{key}']: dim {sym} is {size}, expected {want} (shape {tuple(t.shape)})")
    if "C" in known and "K" in known and known["C"] != known["K"] + 1:
        raise ValueError(f"{schema}: C = {known['C']} classes but K = {known['K']} (K must be C - 1)")
    return known
