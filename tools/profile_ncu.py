"""Single-GPU ncu target: 2 emulated ranks on cuda:0 run ONE allreduce.

ncu serialises kernels, so rank 0's kernel spins until TL_NVL_TIMEOUT (kept short here) and gives up; rank 1's kernel
then finds every flag already raised and runs at full speed - that second launch is the one to profile
(--launch-skip 1 --launch-count 1).  "Peer" traffic stays inside one GPU's HBM here, so the capture is about the
instruction stream / occupancy / memory pipeline of the kernel, not about NVLink."""
import os
import sys

os.environ.setdefault("UCC_TL_NVL_TIMEOUT", "50ms")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ucc_b200 import capi as U  # noqa: E402
from ucc_b200.harness import UccJob, coll_args  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "direct"
nbytes = {"direct": 256 << 20, "staged": 64 << 20, "oneshot": 256 << 10}[mode]
env = {"UCC_TL_NVL_ZCOPY": "y" if mode == "direct" else "n", "UCC_TL_NVL_ZCOPY_THRESH": "0", "UCC_TL_NVL_MAX_BLOCKS": "128"}
torch.cuda.set_device(0)
n = 2
job = UccJob(n, env=env)
team = job.create_team(range(n))
cnt = nbytes // 4
src = [torch.full((cnt,), float(r + 1), device="cuda") for r in range(n)]
dst = [torch.zeros(cnt, device="cuda") for _ in range(n)]
torch.cuda.synchronize()
args = [coll_args("allreduce", dt="float32", mem_type=U.UCC_MEMORY_TYPE_CUDA, src_ptr=src[r].data_ptr(), dst_ptr=dst[r].data_ptr(), count_src=cnt, count_dst=cnt) for r in range(n)]
req = team.coll(args)
req.post()
try:
    st = req.wait(max_iters=50_000_000)
except Exception as e:  # noqa: BLE001
    st = repr(e)
torch.cuda.synchronize()
print("status", st, "dst[1][0] =", float(dst[1][0]), "(expected 3.0)")
os._exit(0)
