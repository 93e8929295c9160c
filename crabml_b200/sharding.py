"""Shard plan of the row/column-sharded decode path and the GGUF byte slicing that goes with it.

The reference is single-device; its unit of parallelism is the output row of matmul_vec (matmul_vec.rs:41-76 deals rows
to the thread pool).  Across N GPUs the same unit gives (Megatron pairing, in GGUF terms):

  rows   (m) of wq / wk / wv    whole heads, contiguous ranges          -> no exchange, KV cache stays local
  cols   (k) of wo              the columns produced by those heads     -> allreduce(sum) of a [dim] f32 row
  rows   (m) of ffn_gate / up   block-aligned share of hidden_dim       -> no exchange
  cols   (k) of ffn_down        the same share, at BLOCK boundaries     -> allreduce(sum) of a [dim] f32 row
  rows   (m) of the classifier  vocab / N                               -> allgather of the logit slices
  token_embd, norm weights      replicated

Column splits must land on quant-block boundaries (32 elements, 256 for K-quants); hidden_dim is dealt in those units, the
first `remainder` ranks take one unit more (11008 = 43 super-blocks -> 6,6,6,5,5,5,5,5 at N = 8).
Pure host logic: numpy only, no torch, no GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# GGUF (bytes per block, elements per block), SURVEY Appendix A
BLOCK = {0: (4, 1), 1: (2, 1), 2: (18, 32), 3: (20, 32), 6: (22, 32), 7: (24, 32), 8: (34, 32), 9: (36, 32),
         10: (84, 256), 11: (110, 256), 12: (144, 256), 13: (176, 256), 14: (210, 256), 15: (292, 256)}


def deal(total_units: int, world: int, rank: int):
    """-> (first unit, unit count) of `rank` when `total_units` are dealt to `world` ranks as evenly as possible."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} of {world}")
    base, rem = divmod(total_units, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class ShardPlan:
    rank: int
    world: int
    q_rows: tuple        # (row0, nrows) of wq                      == columns of wo
    kv_rows: tuple       # (row0, nrows) of wk / wv
    hidden: tuple        # (first, count) in ELEMENTS: rows of gate/up == columns of down
    vocab_rows: tuple    # (row0, nrows) of the classifier

    @property
    def hidden_local(self):
        return self.hidden[1]


def make_plan(n_heads: int, n_kv_heads: int, embedding_dim: int, hidden_dim: int, vocab_size: int, down_dtype: int,
              rank: int, world: int, f16_kv: bool = False) -> ShardPlan:
    if world < 1 or world > 8:
        raise ValueError("world size must be 1..8")
    if n_heads % world or n_kv_heads % world:
        raise ValueError(f"{n_heads} heads / {n_kv_heads} kv heads do not divide by {world}")
    if vocab_size % world or (vocab_size // world) % 4:
        raise ValueError(f"vocab {vocab_size} must divide by {world} into multiples of 4")
    if n_heads != n_kv_heads and not f16_kv and world > 1:
        # F32 cache: query head h reads kv head h % n_kv (batch_matmul.rs:47-71) -- not local for contiguous head ranges
        raise ValueError("grouped-query models shard only with the f16 kv cache")
    hd = embedding_dim // n_heads
    nh, nkv = n_heads // world, n_kv_heads // world
    be = BLOCK[down_dtype][1]
    if hidden_dim % be:
        raise ValueError(f"hidden_dim {hidden_dim} is not a multiple of the {be}-element block")
    if (nh * hd) % be:
        raise ValueError(f"{nh} heads x {hd} is not a multiple of the {be}-element block (wo column split)")
    u0, un = deal(hidden_dim // be, world, rank)
    if un == 0:
        raise ValueError("more ranks than hidden blocks")
    v = vocab_size // world
    return ShardPlan(rank, world, (rank * nh * hd, nh * hd), (rank * nkv * hd, nkv * hd), (u0 * be, un * be), (rank * v, v))


def slice_rows(data: np.ndarray, rows: int, cols: int, dtype: int, row0: int, nrows: int) -> np.ndarray:
    """GGUF bytes of rows [row0, row0+nrows) of a [rows, cols] tensor."""
    bb, be = BLOCK[dtype]
    rb = cols // be * bb
    a = np.asarray(data).view(np.uint8).reshape(-1)[: rows * rb].reshape(rows, rb)
    return np.ascontiguousarray(a[row0:row0 + nrows]).reshape(-1)


def slice_cols(data: np.ndarray, rows: int, cols: int, dtype: int, col0: int, ncols: int) -> np.ndarray:
    """GGUF bytes of columns [col0, col0+ncols) (block aligned) of a [rows, cols] tensor."""
    bb, be = BLOCK[dtype]
    if col0 % be or ncols % be:
        raise ValueError(f"column range [{col0}, +{ncols}) is not aligned to the {be}-element block")
    rb = cols // be * bb
    a = np.asarray(data).view(np.uint8).reshape(-1)[: rows * rb].reshape(rows, rb)
    return np.ascontiguousarray(a[:, col0 // be * bb:(col0 + ncols) // be * bb]).reshape(-1)


# which way each weight of a layer is cut: name -> ("rows"|"cols", plan attribute)
CUTS = {"wq": ("rows", "q_rows"), "wk": ("rows", "kv_rows"), "wv": ("rows", "kv_rows"), "wo": ("cols", "q_rows"),
        "ffn_gate": ("rows", "hidden"), "ffn_up": ("rows", "hidden"), "ffn_down": ("cols", "hidden"),
        "output_weight": ("rows", "vocab_rows")}


def shard_bytes(name: str, data: np.ndarray, rows: int, cols: int, dtype: int, plan: ShardPlan):
    """-> (bytes, [rows, cols]) of this rank's shard of weight `name` (replicated tensors come back whole)."""
    if name not in CUTS or plan.world == 1:
        return np.asarray(data).view(np.uint8).reshape(-1), [rows, cols]
    how, attr = CUTS[name]
    first, count = getattr(plan, attr)
    if how == "rows":
        return slice_rows(data, rows, cols, dtype, first, count), [count, cols]
    return slice_cols(data, rows, cols, dtype, first, count), [rows, count]
