"""torchrun -n 2: distributed 2x2 chain vs the single-rank chain on the same H (run on >= 2 GPUs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from quantool_b200 import cabi
from quantool_b200.engine import pipeline, schemes
def timeit(fn, n=2):
    fn(); torch.cuda.synchronize(); dist.barrier(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for K in (8192 + 128, 14336):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((16384, K), device=dev, dtype=torch.bfloat16, generator=g); x[:, :30] *= 20
    H = torch.zeros((K, K), device=dev); cabi.hessian_accumulate(x, H); cabi.hessian_finalize(H, 2/128); del x
    args = schemes.resolve("W4A16", "group")
    lq = pipeline.GPTQLayerQuantizer(args, dist=pipeline.Dist())
    ctx = lq.prepare_input_distributed(H, slot="#d")
    Ud = ctx.U.clone()
    ref = lq.prepare_input(H, owner=0, slot="#s")
    Us = ref.U
    rel = (torch.linalg.norm(Ud - Us) / torch.linalg.norm(Us)).item()
    lower = torch.tril(Ud, -1).abs().max().item()
    t_d = timeit(lambda: lq.prepare_input_distributed(H, slot="#d"))
    t_s = timeit(lambda: lq.prepare_input(H, owner=0, slot="#s"))
    if rank == 0:
        print(f"K={K} world={dist.get_world_size()} rel diff dist vs single {rel:.3e} lower-max {lower:.1e} info {int(ctx.info.item())} "
              f"time dist {t_d:.1f} ms single(+bcast) {t_s:.1f} ms", flush=True)
    assert rel < 1e-4, rel
    del H, Ud
    lq.drop_scratch(); torch.cuda.empty_cache()
dist.destroy_process_group()
