"""Launch the dominant kernel (weight-streaming GEMM, SiLU epilogue) at the 70B gate|up shape a few times — target of
`ncu --set full -k regex:gemm_ws` (see tools/gpu_ncu.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssd_b200 import ops  # noqa: E402

N, K, M = 28672, 8192, 7   # ffn, hidden of Llama-3.1-70B; gate|up packed = [2*28672, 8192]
w = (torch.randn(2 * N, K, device="cuda") * 0.02).to(torch.bfloat16)
x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
wq = (torch.randn(10240, K, device="cuda") * 0.02).to(torch.bfloat16)   # qkv (split-K partial epilogue path via ops.linear)
for _ in range(5):
    ops.gate_up_silu(x, w)
    ops.linear(x, wq)
torch.cuda.synchronize()
print("done")
