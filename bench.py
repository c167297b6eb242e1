#!/usr/bin/env python
"""bench.py -- MMinterest training step throughput (BASELINE.json metric: train interactions/s).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path (default workload c2)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU PyTorch path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # weak scaling: per-GPU batch fixed

One JSON line on stdout (rank 0).  A "step" = gather+pad+mask+L1-normalise -> forward -> focal loss
-> backward -> (gradient all-reduce) -> clip + AdamW over one batch of synthetic interactions.
Both arms run the reference's TRAINING semantics by default: nn.Dropout(0.1) live at every site (--dropout 0.1; the
driver never overrides SegFormerX's default).  `dropout_off` in the JSON line is the same step with every site off
(the eval()-mode arithmetic the bit-level parity tests use); --dropout 0 makes that the headline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from segmminterest_b200 import synth  # noqa: E402


N_USERS, N_ITEMS = 1903, 352494      # the reference reader's table sizes (utils/dataloader_SegMM.py:78-79)


def model_args(precision, model="image"):
    """image: the single image backbone + focal (BCE family) the metric is quoted on;
    both: the reference's DEFAULT configuration (main...SegMM.py:506-509,527): image backbone + ID backbone fused by
    InteractionAggregation(2 heads), interestBPR, embedding tables of the full dataset's size."""
    it = {"user": "both", "photo": "both"} if model == "both" else {"user": "image", "photo": "image"}
    return SimpleNamespace(debug=0, input_type=it, d_model=512, nhead=16,
                           learnable_bias=0, exposure_prob=[1.0] * 40, fusion_heads=2,
                           loss_type_list=["interestBPR"] if model == "both" else ["focal"],
                           loss_weight={"focal": 1.0, "interestBPR": 1.0}, mask_loss=0, num_layers_enc=6, ablation_type="ours", use_pe=1,
                           mmi_precision=precision)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu),
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_step_builder(wl, b_cpu, seed=0, threads=None, dropout=0.0):
    """The reference's training step on host cores, restated by oracle/mmi_oracle.py (the reference
    is pure Python and does not travel to the GPU box; the port is pinned to it by tests/golden)."""
    from oracle import gather_oracle, mmi_oracle
    from segmminterest_b200.model import reference_state_shapes
    torch.set_num_threads(threads or os.cpu_count())
    shapes = reference_state_shapes(512, 6, wl.din, wl.lt, 40)
    sd = {k: torch.from_numpy(v) for k, v in synth.fill_state_dict(shapes, 42).items()}
    live = mmi_oracle.live_param_names(list(sd.keys()), 6)
    for k in live:
        sd[k].requires_grad_(True)
    params = [sd[k] for k in live]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    n_rows = 1 << 14
    table = synth.make_table(n_rows, wl.din, seed=1234)
    usr_idx, vid_idx, gt = synth.make_indices(b_cpu, wl.lt, wl.segs_per_video, n_rows, seed=2025 + seed)
    state = {"step": 0}

    def step():
        u, um = gather_oracle.gather_dense(table, usr_idx)
        c, cm = gather_oracle.gather_dense(table, vid_idx)
        u = torch.from_numpy(gather_oracle.l1_normalise(u))
        c = torch.from_numpy(gather_oracle.l1_normalise(c))
        for p in params:
            p.grad = None
        out = mmi_oracle.forward(sd, u, torch.from_numpy(um), c, torch.from_numpy(cm), torch.from_numpy(gt), nhead=16, num_layers=6,
                                 drop=mmi_oracle.torch_dropout(dropout) if dropout > 0 else None)
        out["loss"].backward()
        state["step"] += 1
        with torch.no_grad():
            mmi_oracle.clip_and_adamw(params, [p.grad for p in params], m, v, state["step"])
        return float(out["loss"].detach())

    return step


def run_cpu_sample(wl, budget_s, steps, warmup, dropout=0.0):
    """Sizes the sample so (steps+warmup) CPU steps fit in ~budget_s; returns (interactions/s, B, cores, ms/step)."""
    probe = cpu_step_builder(wl, 2, dropout=dropout)
    t0 = time.perf_counter(); probe(); t1 = time.perf_counter(); probe(); t2 = time.perf_counter()
    per_inter = max((t2 - t1) / 2, 1e-3)
    b = int(max(1, min(32, budget_s / ((steps + warmup) * per_inter))))
    step = cpu_step_builder(wl, b, dropout=dropout)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    return b / dt, b, torch.get_num_threads(), dt * 1e3


def reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    val, b, cores, ms = run_cpu_sample(wl, args.ref_budget, steps, warmup, dropout=args.dropout)
    dnote = f"dropout {args.dropout} (torch generator)" if args.dropout > 0 else "dropout off"
    line = {"impl": "reference", "metric": "train_interactions_per_s", "value": val, "unit": "interactions/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": wl.name, "cpu_batch": b, "dropout": args.dropout},
            "cpu_baseline": {"value": val, "unit": "interactions/s", "cores": cores, "kind": "port",
                             "sample": f"{b} interactions/step of {wl.name} (oracle port of the reference's eager PyTorch step, fp32, {dnote})"},
            "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def fp32_leg(args, wl, table, dev, batch=64, steps=2, split_tc=True):
    """The strict-parity fp32 mode (the reference's own precision) on the same workload at a reduced batch.
    split_tc (default): GEMMs on the tensor cores as split-bf16 products; False: the FFMA kernels (MMI_FP32_TC=0)."""
    from segmminterest_b200.model import build_model
    from segmminterest_b200.train import TrainStep
    torch.manual_seed(42)
    os.environ["MMI_FP32_TC"] = "1" if split_tc else "0"
    m = build_model(model_args("fp32"), din=wl.din, max_usr_len=wl.lt).to(dev)
    m.train(args.dropout > 0)
    ts = TrainStep(m, table, global_batch=batch, dropout=args.dropout)
    u, v, gt = synth.make_indices(batch, wl.lt, wl.segs_per_video, wl.n_rows, seed=4242)
    u, v, gt = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), torch.from_numpy(gt).to(dev)
    ts.step(u, v, gt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ts.step(u, v, gt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del ts, m
    torch.cuda.empty_cache()
    os.environ.pop("MMI_FP32_TC", None)
    note = ("fp32 mode (logits / gradients within 1e-4 of the reference): GEMMs as six bf16 products of an exact three-term split on "
            "the tensor cores, attention / LayerNorm fp32 SIMT; same workload shapes, reduced batch" if split_tc else
            "fp32 mode with the FFMA GEMM kernels (MMI_FP32_TC=0)")
    return {"value": batch / (ms * 1e-3), "unit": "interactions/s", "ms_per_step": ms, "batch": batch, "note": note}


def loader_leg(dev, n=256):
    """Batch construction, ours vs the reference's Python loader, on BASELINE config 1: the first 256
    train interactions of SegMM_inter_sample.csv as prepared by the reference's own reader (tests/golden/config1.npz).
    ours: HostFrameIndex (vectorised index build) + device gather + pad + mask, synchronised per batch;
    reference: the oracle's restatement of FrameDatasetSeq_SegMM._getitem + DataCollator on the host cores."""
    import io
    import random
    import pandas as pd
    from oracle import gather_oracle
    from segmminterest_b200.loader import DeviceFrameLoader
    path = os.path.join(ROOT, "tests", "golden", "config1.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    df = pd.read_csv(io.BytesIO(z["files/train_his.csv"].tobytes()), sep="\t").reset_index(drop=True).sort_values(by=["user_id", "time_ms"])
    lineid = json.loads(z["files/SegMM_photoidframeid2lineid.json"].tobytes())
    uid = json.loads(z["files/user_input_dict.json"].tobytes())
    u2i, i2i = json.loads(z["files/second_map_user2id.json"].tobytes()), json.loads(z["files/second_map_item2id.json"].tobytes())
    table = np.random.default_rng(meta["table_seed"]).standard_normal((meta["n_rows"], meta["din"]), dtype=np.float32)
    corpus = SimpleNamespace(data_df={"train": df.iloc[:n]}, user_input_dict=uid)
    # reference port
    rows = df.iloc[:n].to_dict("records")
    random.seed(42)
    t0 = time.perf_counter()
    batch = gather_oracle.collate_port([gather_oracle.getitem_port(r, lineid, uid, table, u2i, i2i) for r in rows])
    t_ref = time.perf_counter() - t0
    # ours
    tab_d = torch.from_numpy(table).to(dev)
    ldr = DeviceFrameLoader(corpus, lineid, tab_d, phase="train", batch_size=n, user2id=u2i, item2id=i2i)
    random.seed(42)
    b0 = next(iter(ldr))
    torch.cuda.synchronize()
    same = bool(np.array_equal(b0["user"].cpu().numpy(), batch["user"]) and np.array_equal(b0["photo"].cpu().numpy(), batch["photo"]))
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        random.seed(42)
        b0 = next(iter(ldr))
        torch.cuda.synchronize()
    t_ours = (time.perf_counter() - t0) / reps
    return {"ours_samples_per_s": n / t_ours, "reference_port_samples_per_s": n / t_ref, "sample": f"{n} interactions of config 1 (Lt 100, Din 1024)",
            "batches_identical": same, "note": "one-time parse of the line-id map / CSV columns excluded on both sides"}


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(synth.WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--micro-batch", type=int, default=0, help="gradient-accumulation slice (configs 3 / 4: saved activations of a slice must fit in HBM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds of CPU work the whole run is sized for")
    ap.add_argument("--model", default="image", choices=["image", "both"],
                    help="image: single image backbone + focal (the quoted configuration); both: the reference's default two-tower model")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32-mode and loader legs (multi-GPU scaling runs)")
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="dropout probability of every nn.Dropout site (reference training default 0.1); 0 = eval()-mode arithmetic")
    args = ap.parse_args()
    wl = synth.WORKLOADS[args.workload]
    if args.impl == "reference":
        reference_arm(args, wl)
        return

    import torch.distributed as dist
    from segmminterest_b200 import ops
    from segmminterest_b200.model import build_model
    from segmminterest_b200.profiler import TIMER
    from segmminterest_b200.train import TrainStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    B = args.batch or wl.batch
    strong = args.workload in ("c3", "c4") and not args.batch      # BASELINE: global batch 4 096 / 8 192 split over the GPUs
    if strong:
        B = max(1, wl.batch // max(world, 1))
    Lt = wl.lt

    torch.manual_seed(42)
    both = args.model == "both"
    model = build_model(model_args(args.precision, args.model), din=wl.din, max_usr_len=Lt, n_users=N_USERS, n_items=N_ITEMS).to(dev)
    model.train(args.dropout > 0)   # train(): every nn.Dropout site of the reference is live (counter-based masks, csrc/dropout.cuh)
    g = torch.Generator(device=dev).manual_seed(1234)
    table = torch.randn(wl.n_rows, wl.din, generator=g, device=dev, dtype=torch.float32)
    ts = TrainStep(model, table, lr=1e-3, weight_decay=1e-4, max_norm=None, global_batch=B * world, dropout=args.dropout)
    ts.gather.valid_hint = {"usr": B * Lt, "vid": B * min(wl.segs_per_video, 40)}     # full histories: algorithmic gather bytes
    idg = torch.Generator(device=dev).manual_seed(77 + rank)
    ids = (dict(usr_id=torch.randint(1, N_USERS + 1, (B,), generator=idg, device=dev), vid_id=torch.randint(1, N_ITEMS + 1, (B,), generator=idg, device=dev))
           if both else {})

    n_batches = 4
    host, devb = [], []
    for i in range(n_batches):
        u, v, gt = synth.make_indices(B, Lt, wl.segs_per_video, wl.n_rows, seed=2025 + 97 * rank + i)
        hu, hv, hg = torch.from_numpy(u).pin_memory(), torch.from_numpy(v).pin_memory(), torch.from_numpy(gt).pin_memory()
        host.append((hu, hv, hg))
        devb.append((hu.to(dev), hv.to(dev), hg.to(dev)))
    staging = (torch.empty_like(devb[0][0]), torch.empty_like(devb[0][1]), torch.empty_like(devb[0][2]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    mb = args.micro_batch
    for i in range(W):
        ts.step(*devb[i % n_batches], micro_batch=mb, **ids)
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = ops.LaunchCounter.n
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        scal = ts.step(*devb[i % n_batches], micro_batch=mb, **ids)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.LaunchCounter.n - n0
    clocks = sampler.stop() if rank == 0 else None
    loss_last = float(scal[3].item())
    value = K * B * world / (ms_total * 1e-3)

    # ---- timed region 2: end to end from pinned host buffers, loss read back every step -----------
    ts.step_host(*host[0], staging, micro_batch=mb, **ids)
    barrier()
    e0.record()
    for i in range(K):
        ts.step_host(*host[i % n_batches], staging, micro_batch=mb, **ids)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_val = K * B * world / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---- the same step with every dropout site off (eval()-mode arithmetic of the parity tests) ---------------------------
    drop_off = None
    if args.dropout > 0:
        ts.dropout = 0.0
        ts.step(*devb[0], micro_batch=mb, **ids)
        barrier()
        e0.record()
        for i in range(K):
            ts.step(*devb[i % n_batches], micro_batch=mb, **ids)
        e1.record()
        barrier()
        ms_off = max_over_ranks(e0.elapsed_time(e1))
        drop_off = {"value": K * B * world / (ms_off * 1e-3), "unit": "interactions/s", "ms_per_step": ms_off / K}
        ts.dropout = args.dropout

    # ---- per-kernel CUDA-event pass for the roofline (one category per kernel AND launch shape) ----------------
    TIMER.enabled = True
    TIMER.detail = True
    TIMER.reset()
    barrier()
    for i in range(K):
        ts.step(*devb[i % n_batches], micro_batch=mb, **ids)
    summ = TIMER.summary()
    TIMER.enabled = False
    TIMER.detail = False
    pk = peaks()
    tot_ms = sum(v["ms"] for v in summ.values())
    dom = max(summ, key=lambda k: summ[k]["ms"])
    d = summ[dom]
    traffic_db = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic_db = json.load(open(tp))
    tr = traffic_db.get(dom)
    if dom.startswith("gather"):
        ach = d["work"] / (d["ms"] * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"]}
    else:
        ach = d["work"] / (d["ms"] * 1e-3) / 1e12
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"]}
    roof["traffic"] = tr["dram_bytes_per_launch"] if tr else None
    if tr:
        roof["traffic_source"] = tr.get("source")
    roof["peak_source"] = pk["src"] + (" (sustained bf16, kernel timed inside a long step)" if roof["bound"] == "tensor" else "")
    roof["share_of_step"] = d["ms"] / tot_ms
    roof["avg_launch_ms"] = d["ms"] / d["launches"]
    roof["launches_per_step"] = d["launches"] // K
    if dom.startswith("attn"):
        # second yardstick for the attention kernels: one ex2 per score (the dq / dk,dv pair recomputes P, so each of its
        # launches pays B*H*Lq*Lk of them; the one-launch backward pays them once for 10 * dh FLOPs per score) against the measured MUFU rate of 16 ex2 / clk / SM (tools/micro/pipe_rates.cu)
        per_flop = {"attn_fwd": 1.0 / (4 * 32), "attn_bwd_dq": 1.0 / (6 * 32), "attn_bwd_dkv": 1.0 / (8 * 32), "attn_bwd_all": 1.0 / (10 * 32),
                    "attn_bwd_fused": 1.0 / (10 * 32)}[dom.split(" ")[0]]
        ex2_per_s = d["work"] * per_flop / (d["ms"] * 1e-3)
        sm_hz = ((clocks or {}).get("sm_mhz") or 1800.0) * 1e6
        roof["mufu"] = {"achieved": ex2_per_s / 1e12, "peak": 16 * 148 * sm_hz / 1e12, "unit": "Tex2/s",
                        "frac": ex2_per_s / (16 * 148 * sm_hz), "peak_source": "16 ex2/clk/SM x 148 SMs x median SM clock sampled during the timed region"}
        roof["note"] = ("head dim 32: one ex2 per score against 128 tensor FLOPs, so this kernel is bounded by the MUFU / issue "
                        "pipes (16 ex2/clk/SM), not by the tensor pipe the FLOP count is divided by")

    def family(k):
        return k.split(" ")[0]
    fam = {}
    for k, v in summ.items():
        f = fam.setdefault(family(k), {"ms": 0.0, "launches": 0, "work": 0.0})
        f["ms"] += v["ms"]; f["launches"] += v["launches"]; f["work"] += v["work"]
    breakdown = {k: {"ms_per_step": round(v["ms"] / K, 3), "launches_per_step": v["launches"] // K,
                     "achieved": (round(v["work"] / (v["ms"] * 1e-3) / (1e9 if k == "gather" else 1e12), 2) if v["work"] else None)}
                 for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    # gather roofline is always reported too (HBM-bound stage the north star names)
    if "gather" in fam:
        gsum = fam["gather"]
        gbs = gsum["work"] / (gsum["ms"] * 1e-3) / 1e9
        gather_roof = {"achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"]}
    else:
        gather_roof = None
    # tensor-pipe utilisation of all GEMM launches together (north star: >= 50 %)
    gemm_roof = None
    if "gemm_tc" in fam:
        gs = fam["gemm_tc"]
        tf = gs["work"] / (gs["ms"] * 1e-3) / 1e12
        gemm_roof = {"achieved": tf, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": tf / pk["tf_sust"], "ms_per_step": gs["ms"] / K}

    # ---- whole-step roofline: algorithmic model FLOPs (live compute, valid tokens) / step time / sustained bf16 peak
    fl = synth.model_flops(Lt, min(wl.segs_per_video, 40), wl.din) * (1 if not both else 1)      # image backbone (the ID tower adds < 8 %)
    step_roof = {"flops_per_interaction": fl, "achieved": fl * B * world / (ms_total / K * 1e-3) / 1e12, "peak": pk["tf_sust"] * world,
                 "unit": "TFLOP/s", "bound": "tensor", "note": "SURVEY 8d formula x interactions / step time, against the sustained bf16 peak x GPUs"}
    step_roof["frac"] = step_roof["achieved"] / step_roof["peak"]

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extras["fp32_mode"] = fp32_leg(args, wl, table, dev)
        extras["fp32_ffma_mode"] = fp32_leg(args, wl, table, dev, split_tc=False)
        extras["loader"] = loader_leg(dev)

    line = {"metric": "train_interactions_per_s", "value": value, "unit": "interactions/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": wl.name, "model": ("both: image + ID backbones, InteractionAggregation(2), interestBPR, tables 1904 x 512 / 352495 x 256"
                                                      if both else "image backbone, focal"), "per_gpu_batch": B, "micro_batch": mb or B, "global_batch": B * world, "hist_len": Lt, "cand_pad": 40,
                       "cand_valid": wl.segs_per_video, "din": wl.din, "d_model": 512, "heads": 16, "layers": 6,
                       "table_rows": wl.n_rows, "parallelism": f"dp{world}", "optimizer": "AdamW lr1e-3 wd1e-4, no clip (the reference's clip_grad_norm_ walks an exhausted generator)",
                       "dropout": (f"{args.dropout} at every reference site (train() mode; counter-based masks, realised drop probability "
                                   f"{round(args.dropout * 256) / 256:.4f})" if args.dropout > 0 else "off (eval()-mode arithmetic)"), "l2": "activations/step >> 126 MB L2 (inputs larger than L2, no flush needed)"},
            "e2e": {"value": e2e_val, "unit": "interactions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "gather_roofline": gather_roof, "gemm_roofline": gemm_roof,
            "step_roofline": step_roof,
            "kernel_breakdown": breakdown, "loss_last": loss_last, "use_tc": bool(ts.engine.use_tc), "dropout_off": drop_off}
    line.update(extras)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            val, b, cores, ms = run_cpu_sample(wl, args.cpu_budget, 2, 1, dropout=args.dropout)
            line["cpu_baseline"] = {"value": val, "unit": "interactions/s", "cores": cores, "kind": "port",
                                    "sample": f"{b} interactions/step of {wl.name}, median of 2 steps after 1 warm-up "
                                              f"({ms:.0f} ms/step; oracle port of the reference's eager PyTorch step, fp32, "
                                              + (f"dropout {args.dropout})" if args.dropout > 0 else "dropout off)")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
