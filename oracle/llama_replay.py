"""Replay of crabml-llama2's Llama2Runner over any class implementing the `Tensor` trait mirror.

TEST INFRASTRUCTURE ONLY.  Used with oracle.tensor_ref.OracleTensor to produce reference
logits / generations, and with the CUDA mirror to cross-check the C++ runner.

Follows crabml-llama2/src/llama2.rs:45-281,527-638 (op order verbatim), model.rs:183-633
(GGUF -> config/weights; dims reversed model.rs:474; norm weights dequantized model.rs:267-282),
sampler.rs:109-116 (argmax = last max), tokenizer/tokenizer_llama.rs:38-136.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F32, F16 = 0, 1
ROPE_LLAMA, ROPE_NEOX = 0, 1


@dataclass
class LlamaConfig:                       # model.rs:30-53
    n_heads: int
    n_kv_heads: int
    n_layers: int
    embedding_dim: int
    hidden_dim: int
    seq_len: int
    vocab_size: int
    rms_norm_eps: float
    rope_dim: int | None
    arch: str = "llama"                  # ModelArchitecture, model.rs:21-27: "llama" | "qwen2" | "gemma"

    def head_size(self):
        return self.embedding_dim // self.n_heads


@dataclass
class LlamaWeights:                      # model.rs:55-84 (llama subset)
    token_embed: object
    wq: list
    wk: list
    wv: list
    wo: list
    ffn_gate_weight: list
    ffn_down_weight: list
    ffn_up_weight: list
    rms_att_weight: list
    rms_ffn_weight: list
    rms_final_weight: object
    output_weight: object | None
    bq: list | None = None               # qwen2 (model.rs bq / bk / bv)
    bk: list | None = None
    bv: list | None = None


class GGUFModel:
    """Raw view of a GGUF file through python `gguf` (v2/v3 files only)."""

    def __init__(self, path):
        import gguf
        self.reader = gguf.GGUFReader(path)
        self.tensors = {t.name: t for t in self.reader.tensors}
        f = self.reader.fields

        def u32(k):
            return int(f[k].parts[f[k].data[0]][0])
        arch = bytes(f["general.architecture"].parts[f["general.architecture"].data[0]]).decode()
        assert arch == "llama", arch
        rope_key = f"{arch}.rope.dimension_count"
        self.tokens = [bytes(f["tokenizer.ggml.tokens"].parts[i]).decode("utf-8") for i in f["tokenizer.ggml.tokens"].data]
        self.scores = [float(f["tokenizer.ggml.scores"].parts[i][0]) for i in f["tokenizer.ggml.scores"].data]
        self.bos = u32("tokenizer.ggml.bos_token_id")
        self.eos = u32("tokenizer.ggml.eos_token_id")
        self.conf = LlamaConfig(
            n_heads=u32(f"{arch}.attention.head_count"),
            n_kv_heads=u32(f"{arch}.attention.head_count_kv"),
            n_layers=u32(f"{arch}.block_count"),
            embedding_dim=u32(f"{arch}.embedding_length"),
            hidden_dim=u32(f"{arch}.feed_forward_length"),
            seq_len=u32(f"{arch}.context_length"),
            vocab_size=len(self.tokens),
            rms_norm_eps=float(np.float32(f[f"{arch}.attention.layer_norm_rms_epsilon"].parts[f[f"{arch}.attention.layer_norm_rms_epsilon"].data[0]][0])),
            rope_dim=u32(rope_key) if rope_key in f else None,
        )

    def raw(self, name):
        """-> (bytes ndarray, shape [rows, cols] (reversed GGUF dims, model.rs:474), ggml type id)"""
        t = self.tensors[name]
        shape = [int(d) for d in reversed(t.shape.tolist())]
        return np.ascontiguousarray(t.data).view(np.uint8).reshape(-1), shape, int(t.tensor_type)

    def has(self, name):
        return name in self.tensors


def load_weights(gm: GGUFModel, T, device) -> LlamaWeights:
    """model.rs:199-460: every tensor goes through T.from_cpu with its own dtype (the
    quantized-GPU relaxation of model.rs:817-837); norm weights are dequantized to F32."""

    def load(name):
        raw, shape, typ = gm.raw(name)
        return T.from_cpu(raw, shape, typ, device)

    def load_norm(name):
        raw, shape, typ = gm.raw(name)
        assert typ == F32, "norm weights are F32 in the fixtures (model.rs:267-282 dequantizes otherwise)"
        return T.from_cpu(raw, shape, F32, device)

    L = gm.conf.n_layers
    return LlamaWeights(
        token_embed=load("token_embd.weight"),
        wq=[load(f"blk.{l}.attn_q.weight") for l in range(L)],
        wk=[load(f"blk.{l}.attn_k.weight") for l in range(L)],
        wv=[load(f"blk.{l}.attn_v.weight") for l in range(L)],
        wo=[load(f"blk.{l}.attn_output.weight") for l in range(L)],
        ffn_gate_weight=[load(f"blk.{l}.ffn_gate.weight") for l in range(L)],
        ffn_down_weight=[load(f"blk.{l}.ffn_down.weight") for l in range(L)],
        ffn_up_weight=[load(f"blk.{l}.ffn_up.weight") for l in range(L)],
        rms_att_weight=[load_norm(f"blk.{l}.attn_norm.weight") for l in range(L)],
        rms_ffn_weight=[load_norm(f"blk.{l}.ffn_norm.weight") for l in range(L)],
        rms_final_weight=load_norm("output_norm.weight"),
        output_weight=load("output.weight") if gm.has("output.weight") else None,
    )


class LlamaTokenizer:                    # tokenizer/tokenizer_llama.rs:5-136
    def __init__(self, tokens, scores, bos, eos):
        self.tokens, self.bos, self.eos = tokens, bos, eos
        self.token_ids = {}
        for i, t in enumerate(tokens):   # later duplicates win the HashMap (:22-26)
            self.token_ids[t] = i
        self.scores = scores

    def decode(self, token) -> bytes:    # :38-58
        piece = self.tokens[token]
        pb = piece.encode("utf-8")
        if pb.startswith(b"<0x") and pb.endswith(b">"):
            return bytes([int(piece[3:-1], 16)])
        if piece.startswith("▁"):
            return piece.replace("▁", " ").encode("utf-8")
        return pb

    def encode(self, text, bos, eos, add_prefix_space=True):     # :62-136
        tokens = []
        text = text.replace(" ", "▁")
        if bos:
            tokens.append(self.bos)
        if add_prefix_space and text:
            if "▁" in self.token_ids:
                tokens.append(self.token_ids["▁"])
        for ch in text:
            if ch in self.token_ids:
                tokens.append(self.token_ids[ch])
            else:
                tokens.extend(b + 3 for b in ch.encode("utf-8"))
        while True:
            best_score, best_idx, best_tok = -math.inf, None, None
            for i in range(len(tokens) - 1):
                tok = self.token_ids.get(self.tokens[tokens[i]] + self.tokens[tokens[i + 1]])
                if tok is not None and self.scores[tok] > best_score:
                    best_score, best_idx, best_tok = self.scores[tok], i, tok
            if best_idx is None:
                break
            tokens[best_idx] = best_tok
            del tokens[best_idx + 1]
        if eos:
            tokens.append(self.eos)
        return tokens


def sample_argmax(logits) -> int:        # sampler.rs:109-116: Iterator::max_by keeps the LAST maximum
    logits = np.asarray(logits)
    m = logits.max()
    return int(np.flatnonzero(logits == m)[-1])


class Llama2Runner:
    """llama2.rs:26-211 generic over the tensor class T (T.alloc / T.from_cpu are classmethods)."""

    def __init__(self, T, conf: LlamaConfig, weights: LlamaWeights, device, seq_len, use_f16_kv_cache=False, world=1):
        """world > 1: `weights` are one rank's shards (crabml_b200/sharding.py) and T has all_reduce_sum_inplace /
        all_gather_from -- the same replay the C++ host runs in sharded mode (llama2_runner.cpp)."""
        self.T, self.conf, self.weights, self.device, self.world = T, conf, weights, device, world
        kv_dtype = F16 if use_f16_kv_cache else F32
        self.logits = np.zeros(conf.vocab_size, np.float32)
        nkv = conf.n_kv_heads // world
        self.key_cache = [T.alloc([nkv, seq_len, conf.head_size()], kv_dtype, device).resize(1, 0) for _ in range(conf.n_layers)]
        self.value_cache = [T.alloc([nkv, seq_len, conf.head_size()], kv_dtype, device).resize(1, 0) for _ in range(conf.n_layers)]

    def kv_cache_len(self):
        return self.key_cache[0].shape()[1]

    def prefill(self, prompt_tokens):    # llama2.rs:111-139 (tokens already encoded)
        base_pos = self.kv_cache_len()
        for pos, token in enumerate(prompt_tokens):
            self.forward([token], base_pos + pos)
        token = sample_argmax(self.logits)
        next_pos = self.kv_cache_len()
        assert next_pos == base_pos + len(prompt_tokens)
        return next_pos, prompt_tokens[-1], token

    def generate(self, pos, token, steps, eos=None):   # llama2.rs:141-172; yields token ids
        max_seq = self.conf.seq_len - pos - 1
        max_steps = max_seq if steps is None else min(max_seq, steps - 1)
        yield token
        cur = token
        for p in range(pos, pos + max_steps):
            self.forward([cur], p)
            new_token = sample_argmax(self.logits)
            if eos is not None and new_token == eos:
                return
            cur = new_token
            yield new_token

    def forward(self, tokens, pos):      # llama2.rs:184-211
        T = self.T
        x = {"llama": self.forward_llama, "qwen2": self.forward_qwen2, "gemma": self.forward_gemma}[self.conf.arch](tokens, pos)   # llama2.rs:186-192
        x_final = T.alloc([self.conf.embedding_dim], F32, self.device)
        x_final.copy_rows_from(x, [len(tokens) - 1])
        output_weight = self.weights.output_weight if self.weights.output_weight is not None else self.weights.token_embed
        logits = output_weight.matmul_vec(x_final)
        if self.world > 1:
            logits = T.alloc([self.conf.vocab_size], F32, self.device).all_gather_from(logits)
        self.logits = np.asarray(logits.export(), np.float32)
        return self.logits

    def forward_llama(self, tokens, pos):        # llama2.rs:213-281
        T, conf, w = self.T, self.conf, self.weights
        embed_dim, n_heads, n_kv_heads = conf.embedding_dim, conf.n_heads // self.world, conf.n_kv_heads // self.world
        head_dim = conf.head_size()
        rope_dim = conf.rope_dim if conf.rope_dim is not None else head_dim
        n_batch = len(tokens)

        x = T.alloc([n_batch, embed_dim], F32, self.device)
        x.copy_rows_from(w.token_embed, tokens)

        for l in range(conf.n_layers):
            x_attn_orig = x.dup()
            x = x.rms_norm_inplace(conf.rms_norm_eps)
            x = x.mul_inplace(w.rms_att_weight[l])
            x = x.with_name(f"attn_rmsnorm:{l}:{pos}")
            x = x.with_name(f"x_debug:{l}:{pos}")

            q = w.wq[l].matmul_vec(x)
            k = w.wk[l].matmul_vec(x)
            v = w.wv[l].matmul_vec(x)

            q = q.reshape([n_batch, n_heads, head_dim])
            k = k.reshape([n_batch, n_kv_heads, head_dim])
            q = q.rope_inplace(ROPE_LLAMA, pos, rope_dim)
            k = k.rope_inplace(ROPE_LLAMA, pos, rope_dim)

            x = self.forward_multi_query_attention(q, k, v, l, pos, n_kv_heads, n_heads, n_heads * head_dim, head_dim, n_batch)
            if self.world > 1:
                x = x.all_reduce_sum_inplace()
            x = x.with_name(f"attn_out:{l}:{pos}")
            x = x.add_inplace(x_attn_orig)
            x = self.forward_ffn(x, l, pos)
            x = x.with_name(f"ffn_out:{l}:{pos}")

        x = x.rms_norm_inplace(conf.rms_norm_eps)
        x = x.mul_inplace(w.rms_final_weight)
        return x.with_name(f"final_rmsnorm:{pos}")

    def forward_qwen2(self, tokens, pos):        # llama2.rs:283-352
        T, conf, w = self.T, self.conf, self.weights
        embed_dim, n_heads, n_kv_heads, head_dim = conf.embedding_dim, conf.n_heads, conf.n_kv_heads, conf.head_size()
        rope_dim = conf.rope_dim if conf.rope_dim is not None else head_dim
        n_batch = len(tokens)
        x = T.alloc([n_batch, embed_dim], F32, self.device)
        x.copy_rows_from(w.token_embed, tokens)
        for l in range(conf.n_layers):
            x_attn_orig = x.dup()
            x = x.rms_norm_inplace(conf.rms_norm_eps)
            x = x.mul_inplace(w.rms_att_weight[l])
            x = x.with_name(f"attn_rmsnorm:{l}:{pos}")
            q = w.wq[l].matmul_vec(x)
            k = w.wk[l].matmul_vec(x)
            v = w.wv[l].matmul_vec(x)
            q = q.add_inplace(w.bq[l])
            k = k.add_inplace(w.bk[l])
            v = v.add_inplace(w.bv[l])
            q = q.reshape([n_batch, n_heads, head_dim])
            k = k.reshape([n_batch, n_kv_heads, head_dim])
            q = q.rope_inplace(ROPE_NEOX, pos, rope_dim)
            k = k.rope_inplace(ROPE_NEOX, pos, rope_dim)
            x = self.forward_multi_query_attention(q, k, v, l, pos, n_kv_heads, n_heads, embed_dim, head_dim, n_batch)
            x = x.with_name(f"attn_out:{l}:{pos}")
            x = x.add_inplace(x_attn_orig)
            x = self.forward_ffn(x, l, pos)
            x = x.with_name(f"ffn_out:{l}:{pos}")
        x = x.rms_norm_inplace(conf.rms_norm_eps)
        x = x.mul_inplace(w.rms_final_weight)
        return x.with_name(f"final_rmsnorm:{pos}")

    def forward_gemma(self, tokens, pos):        # llama2.rs:455-524
        T, conf, w = self.T, self.conf, self.weights
        embed_dim, n_heads, n_kv_heads, head_dim = conf.embedding_dim, conf.n_heads, conf.n_kv_heads, conf.head_size()
        rope_dim = conf.rope_dim if conf.rope_dim is not None else head_dim
        n_batch = len(tokens)
        x = T.alloc([n_batch, embed_dim], F32, self.device)
        x.copy_rows_from(w.token_embed, tokens)
        x = x.scale_inplace(float(np.sqrt(np.float32(embed_dim))))      # (embed_dim as f32).sqrt()
        x = x.with_name("scaled_embed")
        for l in range(conf.n_layers):
            x_attn_orig = x.dup()
            x = x.rms_norm_inplace(conf.rms_norm_eps)
            x = x.mul_inplace(w.rms_att_weight[l])
            x = x.with_name(f"attn_rmsnorm:{l}:{pos}")
            q = w.wq[l].matmul_vec(x)
            k = w.wk[l].matmul_vec(x)
            v = w.wv[l].matmul_vec(x)
            q = q.reshape([n_heads, head_dim])
            k = k.reshape([n_kv_heads, head_dim])
            q = q.rope_inplace(ROPE_NEOX, pos, rope_dim)
            k = k.rope_inplace(ROPE_NEOX, pos, rope_dim)
            x = self.forward_multi_query_attention(q, k, v, l, pos, n_kv_heads, n_heads, embed_dim, head_dim, n_batch)
            x = x.add_inplace(x_attn_orig)
            x = self.forward_ffn(x, l, pos, gelu=True)
            x = x.with_name(f"ffn_out:{l}:{pos}")
        x = x.rms_norm_inplace(conf.rms_norm_eps)
        x = x.mul_inplace(w.rms_final_weight)
        return x.with_name(f"final_rmsnorm:{pos}")

    def forward_multi_query_attention(self, q, k, v, l, pos, n_kv_heads, n_heads, embed_dim, head_dim, n_batch):
        # llama2.rs:527-603
        k = k.reshape([n_batch, n_kv_heads, head_dim]).transpose([1, 0, 2])
        v = v.reshape([n_batch, n_kv_heads, head_dim]).transpose([1, 0, 2])
        self.key_cache[l].concatenate(k, 1)
        self.value_cache[l].concatenate(v, 1)

        q = q.reshape([n_batch, n_heads, head_dim]).transpose([1, 0, 2]).contiguous().scale_inplace(
            float(np.float32(1.0) / np.sqrt(np.float32(head_dim))))

        k_cache = self.key_cache[l]
        k_strider_orig = k_cache.strider().clone()
        k_cache = k_cache.transpose([0, 2, 1])
        attn = q.batch_matmul(k_cache)
        attn = attn.softmax_inplace(2)
        self.key_cache[l] = k_cache.with_strider(k_strider_orig)

        v_cache = self.value_cache[l]
        v_strider_orig = v_cache.strider().clone()
        x_with_attn = attn.batch_matmul(v_cache)
        if n_batch == 1:
            x_with_attn = x_with_attn.reshape([n_batch, embed_dim])
        else:
            x_with_attn = x_with_attn.transpose([1, 0, 2]).contiguous().reshape([n_batch, embed_dim])
        self.value_cache[l] = v_cache.with_strider(v_strider_orig)
        return self.weights.wo[l].matmul_vec(x_with_attn)

    def forward_ffn(self, x, l, pos, gelu=False):    # llama2.rs:605-638 (Activation::SiLU | GeLU)
        w = self.weights
        x_orig_ffn = x.dup()
        x = x.rms_norm_inplace(1e-5)     # literal, llama2.rs:611 (B5)
        x = x.mul_inplace(w.rms_ffn_weight[l])
        h1 = w.ffn_gate_weight[l].matmul_vec(x)
        h2 = w.ffn_up_weight[l].matmul_vec(x)
        h1 = h1.gelu_inplace() if gelu else h1.silu_inplace()
        h1 = h1.mul_inplace(h2)
        x = w.ffn_down_weight[l].matmul_vec(h1)
        if self.world > 1:
            x = x.all_reduce_sum_inplace()
        x = x.add_inplace(x_orig_ffn)
        return x


def decode_text(tok: LlamaTokenizer, ids) -> str:
    return b"".join(tok.decode(i) for i in ids).decode("utf-8", errors="replace")
