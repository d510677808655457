"""temp > 0 parity with a STATED KL tolerance at the full vocabulary (north_star: "within stated KL tolerance at temp>0";
VERDICT r1 'next round' 1e).

The fused kernels draw from an explicit Philox stream, not torch's global generator, so at temp > 0 they cannot be
compared with the reference draw by draw; what has to hold is that they sample from the reference's DISTRIBUTIONS:

  * `Sampler.forward` (ssd/layers/sampler.py:18-36): token ~ softmax(fp32(logits) / T);
  * `verify()` ratio path (ssd/utils/verify.py:107-164): accept x_j while u <= min(1, p(x_j) / (q(x_j) + 1e-10)); on the first
    rejection resample from max(0, p - q) renormalised — by the speculative-sampling theorem the first emitted token is
    then distributed exactly as p = softmax(logits_p / T), whatever q is.

Protocol (V = 128256, logits ~ N(0, 3^2) in bf16, T = 0.7, N = 32768 draws): the empirical distribution is binned into
the 63 most likely tokens under the reference distribution + one bin for the rest, and

    KL(empirical || reference)  <=  KL_TOL = 0.002 nats.

For a correct sampler 2*N*KL is chi-square with 63 degrees of freedom (mean 63, sd 11.2): the tolerance is 2*N*KL <= 131,
six standard deviations above the mean, i.e. it essentially never fails by chance, while e.g. sampling at T = 0.8 instead
of 0.7 or dropping the residual renormalisation costs > 0.01 nats.  The reference distributions are evaluated by the
oracle's `_softmax_rows` (itself pinned against the reference's own tensors in tests/test_oracle_golden.py).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

V = 128256
T = 0.7
N_ROWS = 16
N_CALLS = 2048
KL_TOL = 0.002
TOP = 63


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _binned_kl(samples: torch.Tensor, ref: torch.Tensor):
    """KL(emp || ref) in nats over {the TOP most likely tokens under ref} + {everything else}."""
    top = ref.topk(TOP).indices
    slot = torch.full((ref.numel(),), TOP, dtype=torch.int64)
    slot[top] = torch.arange(TOP)
    cnt = torch.bincount(slot[samples.cpu()], minlength=TOP + 1).double()
    emp = cnt / cnt.sum()
    refb = torch.zeros(TOP + 1, dtype=torch.float64)
    refb.index_add_(0, slot, ref.double())
    nz = emp > 0
    return float((emp[nz] * (emp[nz] / refb[nz]).log()).sum()), int(cnt.sum())


def _logits(seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(V, generator=g) * 3).to(torch.bfloat16)


def test_sampler_kl_at_full_vocab(dev):
    from oracle.verify import _softmax_rows
    from ssd_b200 import ops
    row = _logits(21)
    ref = _softmax_rows(row[None, None], torch.tensor([T]))[0, 0]
    logits = row[None].repeat(N_ROWS, 1).contiguous().to(dev)
    temps = torch.full((N_ROWS,), T, device=dev)
    draws = torch.cat([ops.sample(logits, temps, seed=5, step_id=i) for i in range(N_CALLS)])
    kl, n = _binned_kl(draws, ref)
    print(f"[KL] sampler: KL(emp || softmax(l/T)) = {kl:.5f} nats over {n} draws (tolerance {KL_TOL})")
    assert n == N_ROWS * N_CALLS and kl <= KL_TOL


def test_verify_ratio_path_kl_at_full_vocab(dev):
    """Draft token from the Philox sampler, accept / reject + residual resampling by verify_kernel: the first emitted
    token must follow p; the recovery token, given a rejection at position 0, must follow norm(max(0, p - q)); the
    acceptance rate must be sum(min(p, q))."""
    from oracle.verify import _softmax_rows
    from ssd_b200 import ops
    K = 6
    lp_row, lq_row = _logits(31), _logits(32)
    lq_row = (0.95 * lp_row.float() + 0.31 * lq_row.float()).to(torch.bfloat16)  # a draft correlated with the target: accept rate ~0.51
    p = _softmax_rows(lp_row[None, None], torch.tensor([T]))[0, 0]
    q = _softmax_rows(lq_row[None, None], torch.tensor([T]))[0, 0]
    resid = (p - q).clamp(min=0)
    resid = resid / resid.sum()
    accept_rate = float(torch.minimum(p, q).sum())
    lp = lp_row[None, None].repeat(N_ROWS, K + 1, 1).contiguous().to(dev)
    lq = lq_row[None, None].repeat(N_ROWS, K, 1).contiguous().to(dev)
    lq0 = lq[:, 0].contiguous()
    temps = torch.full((N_ROWS,), T, device=dev)
    spec = torch.zeros(N_ROWS, K + 1, dtype=torch.int64, device=dev)
    first, recov, rejected = [], [], []
    for i in range(N_CALLS):
        spec[:, 1] = ops.sample(lq0, temps, seed=9, step_id=i)
        n_acc, rec = ops.verify(lp, lq, spec, temps, temps, None, True, seed=11, step_id=i)
        rej = n_acc == 0
        first.append(torch.where(rej, rec, spec[:, 1]))
        recov.append(rec[rej])
        rejected.append(rej.sum())
    first, recov = torch.cat(first), torch.cat(recov)
    n_total = N_ROWS * N_CALLS
    kl_first, n1 = _binned_kl(first, p)
    kl_rec, n2 = _binned_kl(recov, resid)
    rate = 1.0 - float(torch.stack(rejected).sum()) / n_total
    sd = math.sqrt(accept_rate * (1 - accept_rate) / n_total)
    tol_rec = KL_TOL * n_total / max(n2, 1)  # same chi-square bound, fewer draws
    print(f"[KL] verify: first token KL(emp || p) = {kl_first:.5f} ({n1} draws, tol {KL_TOL}); recovery | reject "
          f"KL(emp || norm(max(0,p-q))) = {kl_rec:.5f} ({n2} draws, tol {tol_rec:.4f}); accept rate {rate:.4f} vs "
          f"sum(min(p,q)) = {accept_rate:.4f} (sd {sd:.4f})")
    assert kl_first <= KL_TOL
    assert kl_rec <= tol_rec
    assert abs(rate - accept_rate) <= 5 * sd + 1e-3
