"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: sample sharding + all-reduce of
the Hessian, row sharding + ragged all-gather, module-wise ownership/broadcast, GGUF tensor placement."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from quantool_b200.engine import pipeline
        from oracle import gptq as og
        d = pipeline.Dist()
        assert d.on and d.world == world and d.rank == rank
        g = torch.Generator().manual_seed(0)
        n, seq, K, N = 6, 64, 128, 72
        X = torch.randn((n, seq, K), generator=g).to(torch.bfloat16)
        W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
        # 1. sample-sharded raw sums + all-reduce == the whole Hessian
        per = pipeline.row_split(n, world)
        s0 = sum(per[:rank])
        xl = X[s0:s0 + per[rank]].reshape(-1, K).float()
        Hraw = xl.t() @ xl
        d.all_reduce_sum(Hraw)
        H = Hraw * (2.0 / n)
        Ho, cnt = og.make_empty_hessian(K), 0
        for b in range(n):
            Ho, cnt = og.accumulate_hessian(X[b:b + 1], Ho, cnt)
        assert (torch.linalg.norm(H - Ho) / torch.linalg.norm(Ho)).item() < 1e-5
        # 2. one owner computes, everyone receives (module-wise Cholesky distribution)
        t = torch.full((4,), float(rank))
        d.broadcast(t, 1)
        assert t.tolist() == [1.0] * 4
        # 3. rows are independent given H: row-sharded GPTQ + ragged all-gather == unsharded GPTQ
        args = og.scheme_weight_args("W4A16")
        sizes = pipeline.row_split(N, world, 16)
        assert sum(sizes) == N and all(s % 16 == 0 for s in sizes[:-1])
        r0 = sum(sizes[:rank])
        _, wq_l, s_l, _, _ = og.quantize_weight(W[r0:r0 + sizes[rank]], Ho, args)
        wq = d.all_gather_rows(wq_l.contiguous(), sizes)
        sc = d.all_gather_rows(s_l.contiguous(), sizes)
        _, wq_full, s_full, _, _ = og.quantize_weight(W, Ho, args)
        assert torch.equal(wq, wq_full) and torch.equal(sc, s_full)
        # 4. min / max reductions used by SmoothQuant
        mn = torch.tensor([float(rank), -float(rank)])
        d.all_reduce_min(mn)
        assert mn.tolist() == [0.0, -float(world - 1)]
        # 5. the Dist API contract the CUDA path relies on: reductions return their tensor, every lane is its own
        #    process group (one NCCL communicator per concurrently processed input), equal row blocks are gathered
        #    straight into the result, H / U move as plain collectives on CPU tensors
        v = d.all_reduce_sum(torch.tensor([1.0 + rank]))
        assert v is not None and v.tolist() == [sum(1.0 + r for r in range(world))]
        l0, l1 = d.lane(0), d.lane(1)
        assert l0.on and l1.on and l0.group is not None and l0.group is not l1.group
        assert d.lane(1).group is l1.group                      # cached, created once
        a = l0.all_reduce_sum(torch.tensor([float(rank)]))
        b = l1.all_reduce_max(torch.tensor([float(rank)]))
        assert a.tolist() == [float(sum(range(world)))] and b.tolist() == [float(world - 1)]
        eq = l1.all_gather_rows(torch.full((2, 3), float(rank)), [2] * world)
        assert eq.shape == (2 * world, 3) and eq[:, 0].tolist() == [float(r) for r in range(world) for _ in range(2)]
        Hs = torch.full((4, 4), 1.0 + rank)
        l0.all_reduce_hessian(Hs)
        assert Hs[0, 0].item() == sum(1.0 + r for r in range(world))
        Us = torch.full((4, 4), float(rank))
        l1.broadcast_upper(Us, 1)
        assert Us[0, 0].item() == 1.0
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"


def test_row_split_and_placement():
    from quantool_b200.engine import pipeline
    from quantool_b200.engine.gguf_file import assign_devices
    assert pipeline.row_split(128, 8) == [16] * 8
    assert pipeline.row_split(10, 4) == [3, 3, 2, 2]
    assert pipeline.row_split(1024, 8, 16) == [128] * 8
    assert pipeline.row_split(40, 4, 16) == [16, 16, 8, 0]
    place = assign_devices([100, 90, 10, 10, 5], 2)
    loads = [sum(s for s, d in zip([100, 90, 10, 10, 5], place) if d == k) for k in range(2)]
    assert abs(loads[0] - loads[1]) <= 10


def test_triangular_area_bounds_and_layer_key_check():
    """Host helpers of the multi-GPU path: the row / column ranges the 2 x 2 distributed chain hands to the ranks are
    128-aligned, cover [0, n) and carry equal triangular work; a checkpoint with tensors the driver would drop raises."""
    import pytest
    import torch
    from quantool_b200.engine.pipeline import GPTQLayerQuantizer, _check_layer_keys
    for n, parts in ((14336, 8), (7168, 4), (4096, 2), (14336, 3)):
        for grow in (True, False):
            b = GPTQLayerQuantizer._area_bounds(n, parts, grow)
            assert b[0] == 0 and b[-1] == n and len(b) == parts + 1 and all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % 128 == 0 for x in b[:-1])
            # work of a range: rows weighted by their index (grow) or by what is left to their right (shrink)
            w = [(hi * hi - lo * lo) if grow else ((n - lo) ** 2 - (n - hi) ** 2) for lo, hi in zip(b, b[1:])]
            assert max(w) <= 1.35 * (sum(w) / parts), (n, parts, grow, w)
    ok = {f"model.layers.0.{k}": torch.zeros(1) for k in ("self_attn.q_proj.weight", "input_layernorm.weight",
                                                          "self_attn.rotary_emb.inv_freq")}
    ok["model.norm.weight"] = torch.zeros(1)
    _check_layer_keys(ok)
    with pytest.raises(ValueError, match="unsupported decoder-layer tensors"):
        _check_layer_keys({**ok, "model.layers.0.self_attn.q_proj.bias": torch.zeros(1)})
    with pytest.raises(ValueError, match="unsupported decoder-layer tensors"):
        _check_layer_keys({**ok, "model.layers.1.self_attn.q_norm.weight": torch.zeros(1)})
