"""Two ranks over NCCL == one rank on the same global batch, for the image model and for the reference's
default two-tower model whose embedding-table gradients travel row-sparse.  Needs two GPUs: skipped on a one-GPU box
(run with `gpurun --gpus 2`; the outcome of that run is kept under profiles/)."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _args(model):
    it = {"user": "both", "photo": "both"} if model == "both" else {"user": "image", "photo": "image"}
    return SimpleNamespace(debug=0, input_type=it, d_model=64, nhead=2, learnable_bias=0, exposure_prob=[1.0] * 40, fusion_heads=2,
                           loss_type_list=["focal"], loss_weight={"focal": 1.0}, mask_loss=0, num_layers_enc=3, ablation_type="ours", use_pe=1,
                           mmi_precision="fp32")


def _build(model, dev):
    from segmminterest_b200.model import build_model
    torch.manual_seed(5)
    return build_model(_args(model), din=32, max_usr_len=12, n_users=17, n_items=29).to(dev).eval()       # eval(): no dropout masks


def _batch(B):
    from segmminterest_b200 import synth
    u, v, gt = synth.make_indices(B, 12, 6, 256, seed=3, ragged=True, hist_videos=4, segs_per_video=3)
    rng = np.random.default_rng(9)
    uid = rng.integers(1, 18, size=B)
    vid = rng.integers(1, 30, size=B)
    vid[B // 2] = vid[0]            # the same video on both ranks: its table row receives gradient rows from both
    return u, v, gt, uid, vid


def _step(model_name, dev, rows, world_batch, group=None):
    from segmminterest_b200.train import TrainStep
    model = _build(model_name, dev)
    g = torch.Generator(device=dev).manual_seed(1234)
    table = torch.randn(256, 32, generator=g, device=dev)
    ts = TrainStep(model, table, global_batch=world_batch, process_group=group)
    u, v, gt, uid, vid = _batch(world_batch)
    sl = slice(*rows)
    t = lambda x: torch.from_numpy(x[sl].copy()).to(dev)  # noqa: E731
    ids = dict(usr_id=t(uid), vid_id=t(vid)) if model_name == "both" else {}
    for _ in range(2):
        scal = ts.step(t(u), t(v), t(gt), **ids)
    torch.cuda.synchronize()
    return {k: p.detach().cpu().clone() for k, p in model.state_dict().items()}, float(scal[3]), ts


def _worker(rank, world, port, model_name, B, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        per = B // world
        sd, loss, ts = _step(model_name, dev, (rank * per, (rank + 1) * per), B)
        out.put((rank, {k: v.numpy() for k, v in sd.items()}, loss, bool(ts.sparse_tables), ts.buckets.n_collectives))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("model_name", ["image", "both"])
def test_two_ranks_equal_one_rank(model_name):
    import torch.multiprocessing as mp
    B, world = 8, 2
    ref, ref_loss, _ = _step(model_name, torch.device("cuda", 0), (0, B), B)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, model_name, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, sd, loss, sparse, ncoll in res:
        assert sparse == (model_name == "both") and ncoll >= 1
        for k, v in ref.items():
            a, b = sd[k].astype(np.float64), v.numpy().astype(np.float64)
            assert np.linalg.norm(a - b) <= 1e-5 * max(np.linalg.norm(b), 1e-12) + 1e-9, (rank, k)
    # focal is a sum over interactions / B_global: the two ranks' partial losses add up to the one-rank loss
    assert abs(sum(r[2] for r in res) - ref_loss) < 1e-5 * abs(ref_loss)
