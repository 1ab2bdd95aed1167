"""Linear-layer shapes of the LLaMA / OPT models the reference is run on, for
synthetic workloads, and the layer -> GPU partition the reference gets from accelerate's
`device_map="balanced"` (reference utils.py:43-44): contiguous blocks of decoder layers, no communication.

Shapes are [out_features, in_features] of every nn.Linear except lm_head (quant_wrapper.py:53).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple


@dataclass(frozen=True)
class ModelShape:
    name: str
    n_layers: int
    per_layer: Tuple[Tuple[int, int], ...]  # (out, in) of each Linear in one decoder layer, in module order

    def layer_weights(self) -> int:
        return sum(o * i for o, i in self.per_layer)

    def total_weights(self) -> int:
        return self.n_layers * self.layer_weights()

    def matrices(self, layers: range) -> List[Tuple[int, int]]:
        return [s for _ in layers for s in self.per_layer]


def _llama(name, n_layers, hidden, inter, kv_hidden=None):
    kv = kv_hidden or hidden
    # q, k, v, o, gate, up, down
    return ModelShape(name, n_layers, ((hidden, hidden), (kv, hidden), (kv, hidden), (hidden, hidden), (inter, hidden), (inter, hidden), (hidden, inter)))


def _opt(name, n_layers, hidden, ffn):
    # k, v, q, out, fc1, fc2
    return ModelShape(name, n_layers, ((hidden, hidden),) * 4 + ((ffn, hidden), (hidden, ffn)))


MODELS = {
    "opt-125m": _opt("opt-125m", 12, 768, 3072),
    "opt-6.7b": _opt("opt-6.7b", 32, 4096, 16384),
    "llama-2-7b": _llama("llama-2-7b", 32, 4096, 11008),
    "llama-2-13b": _llama("llama-2-13b", 40, 5120, 13824),
    "llama-2-70b": _llama("llama-2-70b", 80, 8192, 28672, kv_hidden=1024),
}


def partition_layers(n_layers: int, world_size: int, rank: int) -> range:
    """Contiguous block of decoder layers owned by `rank` (balanced: sizes differ by at most one, the
    larger blocks first), mirroring accelerate's balanced placement used at reference utils.py:43-44."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_layers, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def split_rows(rows: int, world_size: int, rank: int) -> range:
    """Row block of a matrix that is split across GPUs (per-tensor granularity, BASELINE.json config 4)."""
    return partition_layers(rows, world_size, rank)
