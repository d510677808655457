"""Weights -> device, and construction of the PairRunner from a Config.

Real checkpoints: HF safetensors are read and packed exactly as ssd/utils/loader.py:186-218 + the weight_loader
callbacks do (q|k|v -> qkv_proj, gate|up -> gate_up_proj; column-parallel shards by output rows, row-parallel by
input columns, embedding / lm_head by vocab rows).  Synthetic directories: generated on the device (synth.py)."""
from __future__ import annotations

import glob
import json
import os

import torch

from . import lib as L
from .runner import ModelSpec, PairRunner


def spec_from_config(hf) -> ModelSpec:
    return ModelSpec(hidden=hf.hidden_size, layers=hf.num_hidden_layers, heads=hf.num_attention_heads,
                     kv_heads=hf.num_key_value_heads, head_dim=hf.head_dim, ffn=hf.intermediate_size,
                     vocab=hf.vocab_size, rms_eps=hf.rms_norm_eps, rope_theta=float(hf.rope_theta),
                     qk_norm=("qwen3" in hf.model_type), tie_embed=bool(hf.tie_word_embeddings),
                     max_pos=hf.max_position_embeddings)


def load_safetensors_weights(path: str, spec: ModelSpec, device, tp_size: int = 1, tp_rank: int = 0) -> dict:
    from safetensors import safe_open

    H, KV, hd = spec.heads // tp_size, spec.kv_heads // tp_size, spec.head_dim
    ffn, Vs, d = spec.ffn // tp_size, spec.vocab // tp_size, spec.hidden
    bf = torch.bfloat16
    w = {"layers": [dict() for _ in range(spec.layers)]}
    for lw in w["layers"]:
        lw["qkv"] = torch.empty((H + 2 * KV) * hd, d, dtype=bf, device=device)
        lw["gate_up"] = torch.empty(2 * ffn, d, dtype=bf, device=device)

    def rows(t, n):  # column-parallel: shard output rows
        return t[tp_rank * n:(tp_rank + 1) * n]

    def cols(t, n):  # row-parallel: shard input columns
        return t[:, tp_rank * n:(tp_rank + 1) * n]

    for file in sorted(glob.glob(os.path.join(path, "*.safetensors"))):
        with safe_open(file, "pt", "cpu") as f:
            for name in f.keys():
                t = f.get_tensor(name).to(bf)
                if name == "model.embed_tokens.weight":
                    w["embed"] = rows(t, Vs).to(device).contiguous()
                elif name == "lm_head.weight":
                    w["lm_head"] = rows(t, Vs).to(device).contiguous()
                elif name == "model.norm.weight":
                    w["final_norm"] = t.to(device)
                elif name.startswith("model.layers."):
                    parts = name.split(".")
                    lw, leaf = w["layers"][int(parts[2])], ".".join(parts[3:])
                    if leaf == "self_attn.q_proj.weight":
                        lw["qkv"][:H * hd] = rows(t, H * hd).to(device)
                    elif leaf == "self_attn.k_proj.weight":
                        lw["qkv"][H * hd:(H + KV) * hd] = rows(t, KV * hd).to(device)
                    elif leaf == "self_attn.v_proj.weight":
                        lw["qkv"][(H + KV) * hd:] = rows(t, KV * hd).to(device)
                    elif leaf == "self_attn.o_proj.weight":
                        lw["o"] = cols(t, H * hd).to(device).contiguous()
                    elif leaf == "mlp.gate_proj.weight":
                        lw["gate_up"][:ffn] = rows(t, ffn).to(device)
                    elif leaf == "mlp.up_proj.weight":
                        lw["gate_up"][ffn:] = rows(t, ffn).to(device)
                    elif leaf == "mlp.down_proj.weight":
                        lw["down"] = cols(t, ffn).to(device).contiguous()
                    elif leaf == "input_layernorm.weight":
                        lw["input_norm"] = t.to(device)
                    elif leaf == "post_attention_layernorm.weight":
                        lw["post_norm"] = t.to(device)
                    elif leaf == "self_attn.q_norm.weight":
                        lw["q_norm"] = t.to(device)
                    elif leaf == "self_attn.k_norm.weight":
                        lw["k_norm"] = t.to(device)
    if "lm_head" not in w:  # tie_word_embeddings (models/llama3.py:321-322)
        w["lm_head"] = w["embed"]
    return w


def shard_packed_weights(w: dict, spec: ModelSpec, tp_size: int, tp_rank: int) -> dict:
    """Slice full packed weights (tp=1 layout) into rank `tp_rank`'s shard with the reference's rules:
    qkv / gate_up column-parallel per sub-matrix (layers/linear.py:116-122,148-162), o / down row-parallel
    (:188-193), embedding / lm_head by vocab rows (embed_head.py:41-47), norms replicated."""
    H, KV, hd, ffn, V = spec.heads, spec.kv_heads, spec.head_dim, spec.ffn, spec.vocab
    h, kv, f, vs = H // tp_size, KV // tp_size, ffn // tp_size, V // tp_size
    r = tp_rank
    out = {"embed": w["embed"][r * vs:(r + 1) * vs].contiguous(), "lm_head": w["lm_head"][r * vs:(r + 1) * vs].contiguous(),
           "final_norm": w["final_norm"], "layers": []}
    for lw in w["layers"]:
        q, k, v = lw["qkv"].split([H * hd, KV * hd, KV * hd], dim=0)
        gate, up = lw["gate_up"].chunk(2, dim=0)
        o = {"input_norm": lw["input_norm"], "post_norm": lw["post_norm"],
             "qkv": torch.cat([q[r * h * hd:(r + 1) * h * hd], k[r * kv * hd:(r + 1) * kv * hd],
                               v[r * kv * hd:(r + 1) * kv * hd]]).contiguous(),
             "o": lw["o"][:, r * h * hd:(r + 1) * h * hd].contiguous(),
             "gate_up": torch.cat([gate[r * f:(r + 1) * f], up[r * f:(r + 1) * f]]).contiguous(),
             "down": lw["down"][:, r * f:(r + 1) * f].contiguous()}
        for k2 in ("q_norm", "k_norm"):
            if k2 in lw:
                o[k2] = lw[k2]
        out["layers"].append(o)
    return out


def load_weights(path: str, spec: ModelSpec, device, tp_size: int = 1, tp_rank: int = 0) -> dict:
    marker = os.path.join(path, "ssd_b200_synthetic.json")
    if os.path.exists(marker):
        from .synth import generate_weights
        with open(marker) as f:
            return generate_weights(spec, json.load(f), device, tp_size, tp_rank)
    if glob.glob(os.path.join(path, "*.safetensors")):
        return load_safetensors_weights(path, spec, device, tp_size, tp_rank)
    raise FileNotFoundError(f"{path}: neither *.safetensors nor ssd_b200_synthetic.json")


class _DraftCfg:
    num_kvcache_blocks = 0


def kv_block_bytes(config, spec: ModelSpec, tp_size: int) -> int:
    return 2 * spec.layers * config.kvcache_block_size * (spec.kv_heads // tp_size) * spec.head_dim * 2


def kv_blocks_for(config, spec: ModelSpec, tp_size: int, share: float, reserved: int = 0) -> int:
    """allocate_kv_cache (engine/model_runner.py:446-476): blocks that fit in share * gpu_memory_utilization * free,
    capped at what max_num_seqs sequences of max_model_len (+ prefix-cache slack) can ever use.  `reserved` = bytes
    already promised to another cache out of the same free-memory snapshot (the reference sizes the draft cache from
    what REMAINS after the target's, draft_runner.py:27)."""
    free, _ = torch.cuda.mem_get_info()
    free = max(0, free - reserved)
    block_bytes = kv_block_bytes(config, spec, tp_size)
    fit = int(free * config.gpu_memory_utilization * share) // block_bytes
    want = max(config.max_num_seqs, 1) * config.max_blocks * 2 + 2
    return max(1, min(fit, want))


def build_runner(config, tp_size: int = 1, tp_rank: int = 0, device=None, finalize: bool = True):
    device = torch.device(device or "cuda:0")
    torch.cuda.set_device(device)
    tspec = spec_from_config(config.hf_config)
    dspec = spec_from_config(config.draft_hf_config) if config.speculate else None
    wt = load_weights(config.model, tspec, device, tp_size, tp_rank)
    wd = load_weights(config.draft, dspec, device) if (dspec is not None and tp_rank == 0) else None
    if dspec is not None and tp_rank != 0:
        dspec = None  # the draft is a replica pinned to rank 0
    nbt = kv_blocks_for(config, tspec, tp_size, 0.8 if dspec else 1.0)
    nbd = kv_blocks_for(config, dspec, 1, 0.75, reserved=nbt * kv_block_bytes(config, tspec, tp_size)) if dspec else None
    config.num_kvcache_blocks = nbt
    draft_cfg = _DraftCfg()
    draft_cfg.num_kvcache_blocks = nbd or nbt
    runner = PairRunner(tspec, dspec, spec_k=config.speculate_k if config.speculate else 0,  # workers keep K although they hold no draft
                        max_batch=max(1, config.max_num_seqs), block_size=config.kvcache_block_size,
                        max_model_len=config.max_model_len, num_blocks_target=nbt, num_blocks_draft=nbd, device=device,
                        use_graph=config.use_cuda_graph, use_pdl=config.use_pdl, jit_speculate=config.jit_speculate,
                        tp_size=tp_size, tp_rank=tp_rank)
    runner.bind_weights(L.TARGET, wt)
    if wd is not None:
        runner.bind_weights(L.DRAFT, wd)
    if finalize:
        runner.finalize()
    return runner, draft_cfg
