"""Multi-GPU numerics of the data-parallel step (needs >= 2 GPUs on the box; skipped otherwise):
  * b200fm_allreduce_f32 (csrc/comm.cu, NVLink peer-memory two-shot all-reduce) vs the exact fp64 mean, ragged sizes, many calls;
  * GradSync on 4M-Tiny: after a 2-rank step every rank holds bit-identical parameters, the averaged gradients equal the gradients
    of ONE process on the concatenated global batch (the DDP contract of run_training_4m.py:512), and the weight-gradient
    GEMMs really wrote into the gradient arena (no bucket copies)."""
import os
import random
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
        for p in (ROOT, os.path.join(ROOT, "ml-4m_b200")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from b200fm.compat import build_mod7_embeddings, create_model
        from b200fm.optim import FusedAdamW, param_groups_like_reference
        from b200fm.parallel import GradSync, _P2PTransport
        from oracle import fourm_oracle as O
        out = {}
        # ---- 1. the raw kernel
        n = 3 * 1000 * 1000 + 4 * 37
        tr = _P2PTransport(n, dev, None, n_ctas=4)
        comm = torch.cuda.Stream()
        worst = 0.0
        for it, (off, ln) in enumerate([(0, n), (4 * 5, 4 * 1001), (1024, 2 * 1000 * 1000), (0, 4), (n - 4 * 33, 4 * 33)]):
            g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
            tr.arena.copy_(torch.randn(n, device=dev, generator=g))
            mine = tr.arena.double().clone()
            alls = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(alls, mine)
            want = (sum(alls) / world)[off:off + ln]
            torch.cuda.synchronize()
            dist.barrier()
            comm.wait_stream(torch.cuda.current_stream())
            tr.all_reduce(off, ln, it + 1, comm)
            comm.synchronize()
            worst = max(worst, float((tr.arena[off:off + ln].double() - want).abs().max()))
            if off > 0:
                assert torch.equal(tr.arena[:off].double(), mine[:off])          # outside the chunk: untouched
        out["kernel_err"] = worst
        dist.barrier()
        tr.close()
        # ---- 2. GradSync on 4M-Tiny, 2 samples per rank
        torch.manual_seed(0)
        enc, dec, info = build_mod7_embeddings()
        model = create_model("fm_tiny_6e_6d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec, modality_info=info).to(dev)
        single = None
        if rank == 0:
            torch.manual_seed(0)
            enc2, dec2, info2 = build_mod7_embeddings()
            single = create_model("fm_tiny_6e_6d_swiglu_nobias", encoder_embeddings=enc2, decoder_embeddings=dec2, modality_info=info2).to(dev)
            single.load_state_dict(model.state_dict())
        sync = GradSync(model, transport="p2p", chunk_mb=8, wait_at_end=False)
        groups = sync.split_param_groups(param_groups_like_reference(model, 0.05))
        opt = FusedAdamW(groups, lr=1e-3, betas=(0.9, 0.95))
        opt.pre_group_hook = sync.wait
        full = O.synthetic_mod7_batch(2 * world, seed=5)
        mine = {m: {k: v[2 * rank:2 * rank + 2].to(dev) for k, v in d.items()} for m, d in full.items()}
        losses = []
        for step in range(3):
            random.seed(step)
            loss, _ = sync({m: dict(d) for m, d in mine.items()}, num_encoder_tokens=128, num_decoder_tokens=128)
            loss.backward()
            if step == 1:
                sync.wait()
                torch.cuda.synchronize()
                out["direct"], out["copied"] = sync.stats["direct"], sync.stats["copied"]
                if rank == 0:
                    random.seed(step)
                    single.load_state_dict(model.state_dict())
                    single.zero_grad(set_to_none=True)
                    l1, _ = single({m: {k: v.to(dev) for k, v in d.items()} for m, d in full.items()}, num_encoder_tokens=128, num_decoder_tokens=128)
                    l1.backward()
                    torch.cuda.synchronize()
                    rels = []
                    for (n_, p), (_, ps) in zip(model.named_parameters(), single.named_parameters()):
                        a, b = p.grad.double(), ps.grad.double()
                        rels.append(float((a - b).norm() / (b.norm() + 1e-30)))
                    out["grad_rel_worst"] = max(rels)
                    out["loss_single"], out["loss_rank0"] = float(l1), float(loss)
            opt.step()
            opt.zero_grad(set_to_none=True)
            losses.append(float(loss))
        lt = torch.tensor(losses, device=dev, dtype=torch.float64)
        alll = [torch.empty_like(lt) for _ in range(world)]
        dist.all_gather(alll, lt)
        out["mean_loss_step1"] = float(sum(a[1] for a in alll) / world)
        out["equal"] = sync.params_equal_across_ranks()
        out["losses"] = losses
        torch.cuda.synchronize()
        dist.barrier()
        sync.close()
        q.put((rank, out))
        dist.destroy_process_group()
    except BaseException as e:          # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put((rank, dict(error=f"{e!r}\n{traceback.format_exc()}")))


def _run_ranks(world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert "error" not in res[r], res[r]["error"]
        assert res[r]["kernel_err"] <= (1e-6 if world == 2 else 1e-5)   # fp32 sum of `world` values in one fixed order, then * 1 / world
        assert res[r]["equal"] is True
        assert res[r]["direct"] > 10 * max(1, res[r]["copied"]), (res[r]["direct"], res[r]["copied"])
    print(f"{world}-rank GradSync:", {k: v for k, v in res[0].items() if k != "losses"}, res[0]["losses"], res[1]["losses"])
    # one process on the global batch == mean over ranks (equal per-modality row counts per sample in the synthetic batch)
    assert abs(res[0]["mean_loss_step1"] - res[0]["loss_single"]) <= 2e-3
    assert res[0]["grad_rel_worst"] <= 5e-2                         # bf16 contractions on different batch splits
    assert res[0]["losses"][2] < res[0]["losses"][0]                # it trains


def test_p2p_allreduce_and_gradsync_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    _run_ranks(2, 29688)


def test_p2p_allreduce_and_gradsync_all_ranks():
    """The same on every GPU of the box (4 or 8 ranks: the W > 2 instantiations of the all-reduce kernel, peer rotation, flag slots)."""
    n = min(8, torch.cuda.device_count())
    if n < 3:
        pytest.skip("needs >= 3 GPUs")
    _run_ranks(n, 29690)
