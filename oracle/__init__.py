"""oracle/ — CPU restatement of the reference's sync speculative-decoding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ssd_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only
as the checker / the timed CPU baseline — never as the product path.

Every function cites the reference file:line it restates (paths relative to the reference
root, tanishqkumar/ssd @ 6ed3022).  The restatement is pinned against golden vectors that
oracle/gen_golden.py produced by importing and running the reference's own Python
(ssd.utils.verify.verify, ssd.layers.*, ssd.models.llama3 / qwen3) in the build container;
see tests/golden/ and tests/test_oracle_golden.py.  Parity status: PINNED for verify(T=0),
verify(T>0) acceptance probabilities / recovery distributions, RMSNorm / RoPE / SiLU*mul,
tiny-model logits and whole sync-SD token traces; the temp>0 *draws* use an explicit Philox
stream (oracle/philox.py) because the reference's global torch RNG stream cannot be replayed
by a fused kernel — those are pinned distributionally, not token-for-token.
"""
