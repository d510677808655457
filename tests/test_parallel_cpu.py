"""CPU (gloo, world_size 2) tests of the N>1 host path: SPMD schedulers stay identical when only rank 0 knows the
verdict and broadcasts it; tensor-parallel weight shards reassemble to the full matrices."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spmd_worker(rank, world, port, out_q):
    import random
    from ssd_b200.engine.scheduler import Scheduler
    from ssd_b200.engine.sequence import Sequence
    from ssd_b200.sampling_params import SamplingParams
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    Sequence.block_size = 16
    cfg = types.SimpleNamespace(max_num_seqs=2, max_num_batched_tokens=4096, max_model_len=512, eos=1, speculate=True,
                                speculate_k=4, kvcache_block_size=16, num_kvcache_blocks=40)
    sch = Scheduler(cfg, draft_cfg=types.SimpleNamespace(num_kvcache_blocks=40))
    rng = random.Random(0)  # same prompts on every rank (SPMD)
    seqs = [Sequence([rng.randrange(2, 90) for _ in range(rng.randrange(4, 30))],
                     SamplingParams(temperature=0.0, max_new_tokens=40, ignore_eos=True)) for _ in range(3)]
    for s in seqs:
        sch.add(s)
    verdict_rng = random.Random(1234 + rank)  # ranks would DISAGREE without the broadcast
    tables = []
    while not sch.is_finished():
        batch, is_prefill = sch.schedule()
        if is_prefill:
            first = torch.tensor([verdict_rng.randrange(2, 90) for _ in batch])
            dist.broadcast(first, src=0)  # ssdk_forward_tokens broadcasts the sampled token from rank 0
            for s, t in zip(batch, first.tolist()):
                s.recovery_token_id = t
                s.num_cached_tokens = s.num_draft_cached_tokens = s.num_prompt_tokens
            continue
        K = 4
        verdict = torch.tensor([[verdict_rng.randrange(0, K + 1), verdict_rng.randrange(2, 90)] +
                                [verdict_rng.randrange(2, 90) for _ in range(K)] for _ in batch])
        dist.broadcast(verdict, src=0)  # the in-graph ncclBroadcast of (tokens, n_accept, recovery)
        sufs = [[s.recovery_token_id] + verdict[i, 2:2 + int(verdict[i, 0])].tolist() for i, s in enumerate(batch)]
        tables.append([list(s.block_table) for s in batch])
        sch.postprocess_speculate(batch, sufs, verdict[:, 1].tolist())
    out_q.put((rank, [s.token_ids for s in seqs], tables))
    dist.destroy_process_group()


def test_spmd_schedulers_stay_in_lockstep():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_spmd_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r, (t, b)) for r, t, b in (q.get(timeout=120) for _ in range(2)))
    for p in procs:
        p.join(timeout=30)
    assert res[0][0] == res[1][0], "token histories diverged between ranks"
    assert res[0][1] == res[1][1], "block tables diverged between ranks"
    assert all(len(t) > 10 for t in res[0][0])


def test_tp_weight_shards_reassemble():
    from ssd_b200 import synth
    from ssd_b200.runner import ModelSpec
    spec = ModelSpec(hidden=128, layers=1, heads=4, kv_heads=2, head_dim=64, ffn=256, vocab=512)
    meta = {"seed": 3, "alpha": 0.8, "role": "target"}
    full = synth.generate_weights(spec, meta, "cpu", 1, 0)
    parts = [synth.generate_weights(spec, meta, "cpu", 2, r) for r in range(2)]
    # vocab-parallel embedding / lm_head: row shards (embed_head.py:41-47)
    assert torch.equal(torch.cat([p["embed"] for p in parts]), full["embed"])
    assert torch.equal(torch.cat([p["lm_head"] for p in parts]), full["lm_head"])
    # per-rank shapes follow the column / row parallel rules (linear.py:90-95,148-162,188-193)
    lw = parts[0]["layers"][0]
    assert lw["qkv"].shape == ((2 + 2 * 1) * 64, 128) and lw["o"].shape == (128, 2 * 64)
    assert lw["gate_up"].shape == (2 * 128, 128) and lw["down"].shape == (128, 128)


def test_safetensors_loader_packs_and_shards(tmp_path):
    from safetensors.torch import save_file
    from ssd_b200.loader import load_safetensors_weights
    from ssd_b200.runner import ModelSpec
    spec = ModelSpec(hidden=64, layers=1, heads=4, kv_heads=2, head_dim=16, ffn=128, vocab=96)
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16)
    sd = {"model.embed_tokens.weight": r(96, 64), "lm_head.weight": r(96, 64), "model.norm.weight": r(64),
          "model.layers.0.self_attn.q_proj.weight": r(64, 64), "model.layers.0.self_attn.k_proj.weight": r(32, 64),
          "model.layers.0.self_attn.v_proj.weight": r(32, 64), "model.layers.0.self_attn.o_proj.weight": r(64, 64),
          "model.layers.0.mlp.gate_proj.weight": r(128, 64), "model.layers.0.mlp.up_proj.weight": r(128, 64),
          "model.layers.0.mlp.down_proj.weight": r(64, 128), "model.layers.0.input_layernorm.weight": r(64),
          "model.layers.0.post_attention_layernorm.weight": r(64)}
    save_file(sd, str(tmp_path / "model.safetensors"))
    w1 = load_safetensors_weights(str(tmp_path), spec, "cpu")
    lw = w1["layers"][0]
    assert torch.equal(lw["qkv"], torch.cat([sd["model.layers.0.self_attn.q_proj.weight"], sd["model.layers.0.self_attn.k_proj.weight"],
                                            sd["model.layers.0.self_attn.v_proj.weight"]]))
    assert torch.equal(lw["gate_up"], torch.cat([sd["model.layers.0.mlp.gate_proj.weight"], sd["model.layers.0.mlp.up_proj.weight"]]))
    for rank in range(2):
        w = load_safetensors_weights(str(tmp_path), spec, "cpu", 2, rank)["layers"][0]
        q = sd["model.layers.0.self_attn.q_proj.weight"][rank * 32:(rank + 1) * 32]
        k = sd["model.layers.0.self_attn.k_proj.weight"][rank * 16:(rank + 1) * 16]
        v = sd["model.layers.0.self_attn.v_proj.weight"][rank * 16:(rank + 1) * 16]
        assert torch.equal(w["qkv"], torch.cat([q, k, v]))
        assert torch.equal(w["o"], sd["model.layers.0.self_attn.o_proj.weight"][:, rank * 32:(rank + 1) * 32])
        assert torch.equal(w["down"], sd["model.layers.0.mlp.down_proj.weight"][:, rank * 64:(rank + 1) * 64])
        assert torch.equal(w["gate_up"], torch.cat([sd["model.layers.0.mlp.gate_proj.weight"][rank * 64:(rank + 1) * 64],
                                                    sd["model.layers.0.mlp.up_proj.weight"][rank * 64:(rank + 1) * 64]]))
