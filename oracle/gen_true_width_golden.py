"""Generate tests/golden/true_width_<cfg>.npz: the ORACLE's sync-SD steps at the TRUE widths of the BASELINE models.

    python oracle/gen_true_width_golden.py [llama70b llama8b qwen32b]

TEST INFRASTRUCTURE (like everything under oracle/).  Runs in the build container (CPU, minutes per config); the GPU
test `tests/test_true_width_gpu.py` regenerates the very same weights on the device — `ssd_b200.synth` with
`rng="hash"` is a pure integer function of (stream, row, column), bit-identical on CPU and GPU — runs the same prompt
through `ssdk_forward_tokens` / `ssdk_spec_step`, and compares tokens, accept counts, recovery tokens and the logits at
2048 sampled vocabulary columns (+ the top-1 value) of every row with what the oracle computed here.

Why these weights: hidden / ffn / head counts / vocab are the real ones (70B: K = 28672 down-proj, G = 8, 64 heads; Qwen3:
q/k norm, d = 5120, V = 151936), two decoder layers each; o_proj / down_proj are NOT degenerate (switching the attention
or the MLP branch off moves the target's sampled logits by 9 on average, max 64 — 87 % of them leave the comparison
tolerance), while the
bigram construction keeps every greedy decision at a margin of thousands, so the token path is reproducible and a
mismatch is a real numerical bug, not a near-tie.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
OUT = REPO / "tests" / "golden"

# name: (target shape, draft shape, batch, prompt lengths, K, spec steps)
CONFIGS = {
    "llama70b": ("llama-3.1-70b", "llama-3.2-1b", 1, [200], 6, 2),   # 200-token prompt: 4 prefill chunks
    "llama8b": ("llama-3.1-8b", "llama-3.2-1b", 2, [40, 23], 6, 2),
    "qwen32b": ("qwen3-32b", "qwen3-0.6b", 2, [33, 48], 6, 2),
}
N_COLS = 2048
LAYERS = 2
META = {"seed": 3, "alpha": 0.85, "rng": "hash", "draft_mode": "perm", "o_std": 0.004, "down_std": 0.001}


def specs(shape: str):
    from oracle.model import ModelCfg
    from ssd_b200 import synth
    from ssd_b200.runner import ModelSpec
    h, L, H, KV, hd, ffn, V, eps, theta, mtype, tied = synth.SHAPES[shape]
    ms = ModelSpec(hidden=h, layers=LAYERS, heads=H, kv_heads=KV, head_dim=hd, ffn=ffn, vocab=V, rms_eps=eps,
                   rope_theta=theta, qk_norm=(mtype == "qwen3"), max_pos=1024)
    oc = ModelCfg(hidden=h, layers=LAYERS, heads=H, kv_heads=KV, head_dim=hd, ffn=ffn, vocab=V, rms_eps=eps,
                  rope_theta=theta, qk_norm=(mtype == "qwen3"), max_pos=1024)
    return ms, oc


def prompts_for(lens, vocab):
    g = torch.Generator().manual_seed(77)
    return [torch.randint(0, vocab, (n,), generator=g).tolist() for n in lens]


def sample_cols(vocab: int):
    g = torch.Generator().manual_seed(1234)
    return torch.randperm(vocab, generator=g)[:N_COLS].sort().values


def main():
    from oracle.model import OracleModel
    from oracle.spec import SpecSession, contiguous_block_tables
    from ssd_b200 import synth
    names = sys.argv[1:] or list(CONFIGS)
    torch.set_num_threads(8)
    for name in names:
        tshape, dshape, B, plens, K, n_steps = CONFIGS[name]
        t0 = time.time()
        tms, toc = specs(tshape)
        dms, doc = specs(dshape)
        wt = synth.generate_weights(tms, {**META, "role": "target"}, "cpu")
        wd = synth.generate_weights(dms, {**META, "role": "draft"}, "cpu")
        print(f"[{name}] weights {time.time() - t0:.0f}s", flush=True)
        bs, mb = 256, 2
        s = SpecSession(OracleModel(toc, wt, B * mb, bs), OracleModel(doc, wd, B * mb, bs), K, mb)
        bt = contiguous_block_tables(B, mb)
        prompts = prompts_for(plens, toc.vocab)
        cols = sample_cols(toc.vocab)
        out = {"cols": cols.numpy(), "B": np.array(B), "K": np.array(K), "n_steps": np.array(n_steps)}
        for b, p in enumerate(prompts):
            out[f"prompt{b}"] = np.array(p, dtype=np.int64)
        rec = s.prefill(prompts, [0.0] * B, bt, bt.clone())
        out["rec0"] = np.array(rec, dtype=np.int64)
        print(f"[{name}] prefill {time.time() - t0:.0f}s rec0={rec}", flush=True)
        for st in range(n_steps):
            suffixes, rec, lp, lq, spec = s.spec_step()
            out[f"s{st}_spec"] = spec.numpy()
            out[f"s{st}_nacc"] = np.array([len(x) - 1 for x in suffixes], dtype=np.int32)
            out[f"s{st}_rec"] = np.array(rec, dtype=np.int64)
            out[f"s{st}_lp"] = lp[..., cols].contiguous().view(torch.int16).numpy()
            out[f"s{st}_lq"] = lq[..., cols].contiguous().view(torch.int16).numpy()
            out[f"s{st}_lp_top"] = lp.float().max(-1).values.numpy()
            out[f"s{st}_lq_top"] = lq.float().max(-1).values.numpy()
            out[f"s{st}_lp_margin"] = (lp.float().topk(2, -1).values[..., 0] - lp.float().topk(2, -1).values[..., 1]).numpy()
            out[f"s{st}_lq_margin"] = (lq.float().topk(2, -1).values[..., 0] - lq.float().topk(2, -1).values[..., 1]).numpy()
            print(f"[{name}] step {st} {time.time() - t0:.0f}s nacc={out[f's{st}_nacc'].tolist()} "
                  f"min margin p={out[f's{st}_lp_margin'].min():.1f} q={out[f's{st}_lq_margin'].min():.1f} "
                  f"|lp| sampled rms={lp[..., cols].float().pow(2).mean().sqrt():.1f}", flush=True)
        np.savez_compressed(OUT / f"true_width_{name}.npz", **out)
        print(f"[{name}] written ({time.time() - t0:.0f}s)", flush=True)


if __name__ == "__main__":
    main()
