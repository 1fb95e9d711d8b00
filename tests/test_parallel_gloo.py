"""Host logic of b200fm.parallel.GradSync (the DDP replacement of the 4M train step) on CPU: world_size 2, gloo.
The arena layout, the per-step claim / hook / chunk bookkeeping, multi-producer (tied) parameters, unused parameters, no_sync and
the optimizer split are transport-independent; the NVLink all-reduce kernel itself is covered by tests/test_gpu_parallel.py."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.emb = torch.nn.Embedding(50, 16)
        self.fc1 = torch.nn.Linear(16, 40, bias=False)
        self.fc3 = torch.nn.Linear(16, 40, bias=False)
        self.fc2 = torch.nn.Linear(40, 16, bias=False)
        self.norm = torch.nn.LayerNorm(16)
        self.unused = torch.nn.Linear(16, 16)
        self.head = torch.nn.Linear(16, 50, bias=False)
        self.head.weight = self.emb.weight                  # tied: two gradient producers per step

    def forward(self, ids):
        x = self.emb(ids)
        x = x + self.fc2(torch.nn.functional.silu(self.fc1(x)) * self.fc3(x))
        return self.head(self.norm(x))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200fm.parallel import GradSync
    torch.manual_seed(100 + rank)                           # different init per rank: GradSync must broadcast rank 0's
    model = _Net()
    sync = GradSync(model, transport="collective", chunk_mb=0.002, small_numel=64)
    assert len(sync.chunks) >= 2
    ref = _Net()
    ref.load_state_dict({k: v.clone() for k, v in model.state_dict().items()})
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    out = {}
    for step in range(3):
        g = torch.Generator().manual_seed(7 * step)
        ids_all = torch.randint(0, 50, (world, 4, 6), generator=g)
        tgt_all = torch.randint(0, 50, (world, 4, 6), generator=g)
        loss = torch.nn.functional.cross_entropy(sync(ids_all[rank]).flatten(0, 1), tgt_all[rank].flatten())
        loss.backward()
        # single-process reference: mean over the ranks' losses == mean of per-rank gradients
        ref.zero_grad(set_to_none=True)
        rl = sum(torch.nn.functional.cross_entropy(ref(ids_all[r]).flatten(0, 1), tgt_all[r].flatten()) for r in range(world)) / world
        rl.backward()
        for (n, p), (_, rp) in zip(model.named_parameters(), ref.named_parameters()):
            if rp.grad is None:
                assert p.grad is not None and float(p.grad.abs().sum()) == 0.0, n       # unused parameter: mean of zeros (DDP semantics)
                continue
            assert torch.allclose(p.grad, rp.grad, rtol=1e-5, atol=1e-6), (step, n)
            assert p.grad.data_ptr() == sync.arena.data_ptr() + p._b200fm_slot.offset * 4, n   # the gradient LIVES in the arena
        opt.step()
        with torch.no_grad():
            for p, rp in zip(model.parameters(), ref.parameters()):
                if rp.grad is not None:
                    rp -= 0.1 * rp.grad
        opt.zero_grad(set_to_none=True)
    out["equal"] = sync.params_equal_across_ranks()
    # no_sync: gradients stay local
    with sync.no_sync():
        loss = sync(torch.full((2, 3), rank, dtype=torch.long)).sum()
        loss.backward()
    out["local_grad"] = float(model.fc2.weight.grad.abs().sum())
    groups = sync.split_param_groups([dict(params=list(model.parameters()), weight_decay=0.1)])
    out["groups"] = [len(g["params"]) for g in groups]
    out["n_params"] = len(list(model.parameters()))
    out["w"] = model.fc2.weight.detach().flatten().tolist()        # plain python: the worker may exit before the parent reads
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_matches_single_process_mean_gradients_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29677, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert res[r]["equal"] is True
        assert sum(res[r]["groups"]) == res[r]["n_params"] and 2 <= len(res[r]["groups"]) <= 4     # sub-groups along the chunk order
    assert res[0]["w"] == res[1]["w"]
    assert res[0]["local_grad"] != res[1]["local_grad"]                  # no_sync left different local gradients


def test_arena_layout_pairs_swiglu_weights():
    """fc1 / fc3 of a SwiGLU block sit back to back (padded to 8 rows) so their weight gradient is ONE GEMM output."""
    sys.path.insert(0, os.path.join(ROOT, "ml-4m_b200"))
    from b200fm.parallel import GradSync
    m = _Net()
    params = GradSync._ordered_params(m)
    names = [n for n, _ in params]
    i1, i3 = names.index("fc1.weight"), names.index("fc3.weight")
    assert i3 == i1 + 1
    assert len({id(p) for _, p in params}) == len(params)                # the tied table appears once
