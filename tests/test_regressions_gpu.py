"""Regression tests (GPU) for host-side hazards found in review: stale operand caches across CUDA-graph replays,
the loader's write-after-read on its device buffers, and the reduction of the data-parallel gradient all-reduce
when the loss is the global-batch CCC."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import jmt_b200  # noqa: E402
from oracle import jmt_oracle as O  # noqa: E402

DEV = "cuda"


def test_eval_between_graph_replays_sees_current_weights():
    """A replayed graph that contains optimizer.step() changes the parameters without bumping their `_version`: an eager
    (validation) forward between replays must re-cast its bf16 operand copies every time -- replay, eval, replay, eval,
    each eval compared with a FRESH module loaded with the current weights."""
    torch.manual_seed(0)
    B, T = 3, 16
    model = jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "TRANSFORMER", "FC", 512, precision="bf16").to(DEV).train()
    opt = torch.optim.SGD(model.live_parameters(), lr=0.5)
    aud, vis = (t.to(DEV) for t in O.synth_features(B, T, [512, 512], 3))
    lv, la = (t.to(DEV) for t in O.synth_labels(B, T, 4))
    crit = jmt_b200.CCCLoss(digitize_num=1)
    n = B * T

    def step(aud, vis, lv, la):
        v, a = model(aud, vis)
        loss = crit(v.view(-1, n), lv.view(-1, n)) + crit(a.view(-1, n), la.view(-1, n))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    # an eager no-grad forward BEFORE capture publishes cache entries, as a validation epoch would
    model.eval()
    with torch.no_grad():
        model(aud, vis)
    model.train()
    g = jmt_b200.GraphedStep(step, [(aud, vis, lv, la)], warmup=1)
    evals = []
    for rnd in range(3):
        g.replay(0)
        model.eval()
        with torch.no_grad():
            v, a = model(aud, vis)
        fresh = jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "TRANSFORMER", "FC", 512, precision="bf16").to(DEV).eval()
        fresh.load_state_dict(model.state_dict(), strict=True)
        with torch.no_grad():
            vf, af = fresh(aud, vis)
        torch.cuda.synchronize()
        assert torch.equal(v, vf) and torch.equal(a, af), f"eval after replay {rnd} ran with stale weights"
        evals.append(v.clone())
        model.train()
    assert not torch.equal(evals[0], evals[1]) and not torch.equal(evals[1], evals[2]), "the replays must change the weights"


def test_loader_does_not_overwrite_batches_still_in_use(tmp_path):
    """FeatureShardLoader hands out two device buffer sets; the H2D copy of batch i+2 must wait for the consumer's kernels
    that read batch i.  The consumer never synchronises (as graph-replay training does) and queues a long kernel before it
    reads each batch."""
    from jmt_b200 import features as F
    rng = np.random.RandomState(2)
    W, T, B, Cv, Ca = 48, 64, 4, 256, 64
    vis = rng.randn(W, Cv, T).astype(np.float32)
    aud = rng.randn(W, T, Ca).astype(np.float32)
    lv = rng.uniform(-1, 1, (W, T)).astype(np.float32)
    la = rng.uniform(-1, 1, (W, T)).astype(np.float32)
    p = str(tmp_path / "s.jmtshard")
    F.write_shard(p, vis, aud, lv, la)
    busy = torch.randn(4096, 4096, device=DEV)
    sums = []
    for aud_d, vis_d, lv_d, la_d in F.FeatureShardLoader([p], B, DEV):
        for _ in range(6):                                 # ~ms of queued work in front of the read of this batch
            busy = torch.nn.functional.normalize(busy @ busy, dim=1)
        sums.append(torch.stack([vis_d.float().sum(), aud_d.float().sum(), lv_d.sum(), la_d.sum()]))
    got = torch.stack(sums).cpu().double().numpy()
    vb = torch.from_numpy(vis).to(torch.bfloat16).float().numpy()
    ab = torch.from_numpy(aud).to(torch.bfloat16).float().numpy()
    want = np.array([[vb[i:i + B].sum(), ab[i:i + B].sum(), lv[i:i + B].sum(), la[i:i + B].sum()] for i in range(0, W, B)])
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-3, atol=5e-2), np.abs(got - want).max()


def _global_loss_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
        import torch.distributed as dist
        torch.cuda.set_device(0)
        jmt_b200.dist.init_from_env("gloo")                # two processes on ONE GPU: gloo carries the CUDA tensors
        B, T = 4, 12
        torch.manual_seed(1)
        model = jmt_b200.Two_transformers(0.0, 0.0, 2, 1, "TRANSFORMER", "SELF_ATTEN", 512, precision="bf16x3").to(DEV).train()
        jmt_b200.dist.broadcast_parameters(model)
        aud, vis = (t.to(DEV) for t in O.synth_features(B, T, [512, 512], 8))
        lv, la = (t.to(DEV) for t in O.synth_labels(B, T, 9))
        lo, hi = jmt_b200.dist.shard_bounds(B, rank, world)
        model.set_grad_sync(jmt_b200.dist.make_grad_sync(global_loss=True))
        crit = jmt_b200.CCCLoss(digitize_num=1, global_stats=True)
        v, a = model(aud[lo:hi], vis[lo:hi])
        loss = crit(v.reshape(1, -1), lv[lo:hi].reshape(1, -1)) + crit(a.reshape(1, -1), la[lo:hi].reshape(1, -1))
        loss.backward()
        sharded = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
        sharded_loss = float(loss.item())
        # the same global batch in one piece, no collectives
        model.set_grad_sync(None)
        model.zero_grad(set_to_none=True)
        crit1 = jmt_b200.CCCLoss(digitize_num=1)
        v, a = model(aud, vis)
        loss1 = crit1(v.reshape(1, -1), lv.reshape(1, -1)) + crit1(a.reshape(1, -1), la.reshape(1, -1))
        loss1.backward()
        torch.cuda.synchronize()
        worst = 0.0
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            den = float(p.grad.norm()) + 1e-12
            worst = max(worst, float((sharded[k] - p.grad).norm()) / den)
        q.put((rank, worst, abs(sharded_loss - float(loss1.item()))))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:      # noqa: BLE001
        q.put((rank, repr(e), None))


def test_global_ccc_loss_gradients_sum_over_ranks():
    """CCCLoss(global_stats=True) + make_grad_sync(global_loss=True) on two ranks == one process on the concatenated batch
    (parameter gradients and loss); averaging the bucket instead would give half of it."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_global_loss_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
    for rank, worst, dl in res:
        assert isinstance(worst, float), (rank, worst)
        assert worst < 2e-3 and dl < 1e-5, (rank, worst, dl)
