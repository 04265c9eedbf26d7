#!/usr/bin/env python
"""bench.py — attention-decoder train-step throughput (captions/s) on N B200 of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp32|bf16|auto]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): basic_att AttentionDecoder,
A = D = E = 512, V = 9490, batch 512 captions PER GPU, all captions 25 tokens (T = 24 decode steps — the reference's
own training regime), synthetic 14x14x2048 features, random-init weights, embedding frozen,
dropout 0.5 in train mode.  One step = forward + loss (models/attention.py:401-414) + backward + gradient
all-reduce (N > 1) + clamp(+-5) + Adam(1e-4).  Weak scaling: per-GPU work is fixed.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: the same step driven from pinned HOST
buffers through the public module API, H2D copies of features+captions and the D2H read of the loss inside the timed
region.  `roofline`: the fused attention-step forward kernel (HBM-bound), algorithmic bytes / live CUDA-event time.
`cpu_baseline` / `--impl reference`: the reference algorithm (oracle/decoders.py, per-step enc_att recompute as at
models/attention.py:54) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

V, A, D, E, P, C, MAXLEN = 9490, 512, 512, 512, 196, 2048, 25
PROF_STRIDE = 11         # attention-step launches per direction per step = MAXLEN - 1 = 24: stride 11 (coprime) visits every time step
B_PER_GPU = 512
CPU_SAMPLE_B = 32
# fused attention step, forward, fp32 features: P*C*4 + P*A*4 + (A + C + C + P)*4 per (image, step)
# (= 2 026 256 B); with bf16-STORED features (the bf16 tier) the two big streams halve: 1 022 736 B.
def att_bytes_per_row(precision):
    s = 2 if precision == "bf16" else 4
    fwd = P * C * s + P * A * s + (A + C + C + P) * 4
    bwd = P * C * s + P * A * s + (3 * C + 2 * A + 3 * P + C) * 4
    return fwd, bwd


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="icd_b200", choices=["icd_b200", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "fp32x3", "bf16"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="captions per GPU (default: the benchmark config)")
    ap.add_argument("--workload", default="train", choices=["train", "beam", "baseline", "glove"],
                    help="train = BASELINE.json configs[2] (the metric's configuration, default); the others time the "
                         "remaining configs and print their own JSON line")
    ap.add_argument("--beam-images", type=int, default=1024, help="images per GPU for the beam-search measurement")
    ap.add_argument("--no-beam", action="store_true", help="skip the beam-5 side measurement of the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference algorithm on the host cores (oracle port)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, sample_b=CPU_SAMPLE_B):
    """Time `steps` train steps of the reference algorithm on a `sample_b`-caption slice of the benchmark batch."""
    import torch
    from icd_b200 import synthetic
    from icd_b200.vocabulary import synthetic_vocab
    import icd_b200.models.attention as my_att
    from oracle import decoders as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    p = my_att.AttentionDecoderParams()
    p.vocab = synthetic_vocab(V)
    torch.manual_seed(0)
    dec = my_att.AttentionDecoder(torch.device("cpu"), p)       # used as a seeded weight container only
    dec.fine_tune_embeddings(False)
    w = {k: v.detach().clone().requires_grad_(k != "embedding.weight") for k, v in dec.state_dict().items()}
    params = [t for t in w.values() if t.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    enc = synthetic.features(sample_b)
    caps, lens = synthetic.captions(sample_b, V, max_len=MAXLEN)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        dl = [l - 1 for l in lens]
        masks = [(torch.rand(sum(x > t for x in dl), D) >= 0.5).float() for t in range(max(dl))]
        preds, _, dl, alphas = O.attention_decoder_forward(w, enc, caps, lens, dropout_p=0.5, dropout_masks=masks)
        loss = O.attention_loss(preds, caps, dl, alphas)
        opt.zero_grad()
        loss.backward()
        for t in params:
            t.grad.clamp_(-5.0, 5.0)
        opt.step()
        loss.item()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return dict(value=sample_b * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                sample="%d-caption slice of the %d-caption batch, T=24, V=%d, %d timed steps" %
                       (sample_b, B_PER_GPU, V, len(times)))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "attention-decoder train-step captions/s", "value": r["value"],
        "unit": "captions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "fp32", B_PER_GPU),
        "cpu_baseline": {"value": r["value"], "unit": "captions/s", "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n, precision, batch):
    return {"workload": "configs[2]: basic_att attention decoder train step, batch %d/GPU, T=24 (25-token captions), "
                        "V=9490, A=D=E=512, 14x14x2048 synthetic features, embedding frozen, dropout 0.5" % batch,
            "global_batch": batch * n, "per_gpu_batch": batch, "decode_steps": 24, "vocab": V,
            "gemm_precision": precision, "parallelism": "dp%d" % n,
            "l2": "inputs_exceed_l2 (822 MB of features per step vs 126 MB L2; no explicit flush)",
            "loss_gradient": ("bf16 tier: attention_caption_loss(bf16_grad_only=True) - the logit gradient reaches the decoder "
                              "backward as its bf16 copy only (the fp32 copy is not materialised)")
                             if precision == "bf16" else "fp32"}


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region, in-process through NVML (nvidia_ml_py):
    a background thread polls every 10 ms with true timestamps (no measurable effect on the step: 9.00 ms at 50, 10 and 5 ms).  NVML is initialised before warm-up (its start-up
    stalls kernel launches for tens of ms, which must not land inside the timed region).  Falls back to an
    `nvidia-smi -lms` child process when the NVML bindings are unavailable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        self.rows, self.proc, self.nvml, self.stop_flag = [], None, None, False
        self.period = float(os.environ.get("ICD_BENCH_SAMPLER_PERIOD", "0.01"))
        try:
            if os.environ.get("ICD_BENCH_SAMPLER", "on") == "smi":
                raise RuntimeError("forced nvidia-smi sampler")
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.h = pynvml, h
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        names = (("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksEventReasonSwPowerCap),
                 ("hw_power_brake", n.nvmlClocksEventReasonHwPowerBrakeSlowdown))
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1e3
                self.rows.append((time.perf_counter(), sm, self.max_sm, pw, [k for k, bit in names if mask & bit]))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.strip().split(",")]
            try:
                reasons = [name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7])
                           if val.lower().startswith("active")]
                self.rows.append((time.perf_counter(), float(f[0]), float(f[1]), float(f[2]), reasons))
            except Exception:
                continue

    def wait_ready(self, timeout=20.0):
        t0 = time.perf_counter()
        while (self.nvml or self.proc) and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def window(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        if not rows:      # a timed region shorter than one polling period: take the nearest sample
            rows = sorted(self.rows, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:1]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({x for r in rows for x in r[4]})
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
                "power_w_max": max(r[3] for r in rows), "reasons": reasons, "samples": len(rows),
                "source": "nvml" if self.nvml else "nvidia-smi"}

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes read+write per launch of the fused attention-step forward kernel from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("att_step_fwd_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


def beam_measure(dev, n_img, k=5, max_steps=24, reps=2):
    """BASELINE.json configs[4]: beam search k = 5 over synthetic features, caption cap 25 tokens (loop body runs for step
    = 1 .. max_steps + 1), weights from the beam fixture recipe so that captions terminate.  Timed region:
    the whole batched decode (icd_beam_search, fp32 tier) + the device->host read of lengths and token ids."""
    import torch
    from icd_b200.gen_captions import beam_search_batched
    from icd_b200.vocabulary import synthetic_vocab
    import icd_b200.models.attention as my_att
    p = my_att.AttentionDecoderParams()
    p.vocab = synthetic_vocab(V)
    torch.manual_seed(0)
    dec = my_att.AttentionDecoder(dev, p)
    with torch.no_grad():
        dec.embedding.weight *= 30.0
        dec.fc.weight[V - 2] *= 30.0
        dec.fc.bias[V - 2] = -4.0
    dec = dec.to(dev).eval()
    g = torch.Generator().manual_seed(77)
    feats = torch.randn(n_img, 14, 14, C, generator=g).abs_().to(dev)
    times, res = [], None
    with torch.no_grad():
        for i in range(reps + 1):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = beam_search_batched(dec, feats, k, V - 3, V - 2, max_steps=max_steps, want_alphas=False)
            lens = res["len"].cpu()
            seqs = res["seq"].cpu()
            e1.record()
            torch.cuda.synchronize()
            if i > 0:
                times.append(e0.elapsed_time(e1))
    ms = min(times)
    done = int((lens > 0).sum())
    return {"metric": "beam-5 captions/s", "value": n_img / (ms / 1e3), "unit": "captions/s", "images": n_img, "beam": k,
            "max_caption_tokens": max_steps + 1, "ms": ms, "completed": done,
            "precision": "fp32x3 (3-term bf16 split on tcgen05, fp32-grade logits; captions identical to the reference on the goldens)",
            "mean_len": float(lens[lens > 0].float().mean()) if done else 0.0}


def side_workload(args):
    """configs[1] (baseline LSTM decoder B=128, L=25), configs[3] (glove_att: E=300, fp64 fine-tuned table, B=512) and
    configs[4] (beam search) on one GPU: fwd + loss + bwd + clamp + Adam, CUDA-event timed."""
    import torch
    import __graft_entry__
    __graft_entry__.build()
    from icd_b200 import synthetic
    from icd_b200.losses import attention_caption_loss, baseline_caption_loss
    from icd_b200.parallel import DataParallelClipAdam
    from icd_b200.vocabulary import synthetic_vocab
    import icd_b200.models.attention as my_att
    import icd_b200.models.baseline as my_base
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if args.workload == "beam":
        line = beam_measure(dev, args.beam_images)
        line.update({"n_gpus": 1, "data": "synthetic", "config": {"workload": "configs[4]: beam search k=5, %d images" % args.beam_images}})
        print(json.dumps(line), flush=True)
        return
    if args.workload == "baseline":
        Bb, L = 128, MAXLEN
        p = my_base.BaselineDecoderParams()
        p.vocab_size = V
        torch.manual_seed(0)
        dec = my_base.BaselineDecoder(p).to(dev)
        dec.precision = "bf16" if args.precision in ("auto", "bf16") else args.precision
        opt = DataParallelClipAdam(dec, lr=1e-4, grad_clip=5.0)
        img = torch.randn(Bb, E, device=dev, requires_grad=True)
        caps, _ = synthetic.captions(Bb, V, max_len=L)
        caps = caps.to(dev)

        def step():
            out = dec(img, caps)
            loss = baseline_caption_loss(out, caps)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        name, batch = "configs[1]: baseline LSTM decoder, batch 128, L=25, V=9490, E=H=512", Bb
    else:
        Bb = args.batch
        p = my_att.AttentionDecoderParams()
        p.vocab = synthetic_vocab(V)
        p.embed_size = 300
        torch.manual_seed(0)
        dec = my_att.AttentionDecoder(dev, p)
        dec.load_pretrained_embeddins(synthetic.glove_like_table(V, 300))
        dec.fine_tune_embeddings(True)
        dec = dec.to(dev)
        dec.precision = "bf16" if args.precision in ("auto", "bf16") else args.precision
        dec.train()
        opt = DataParallelClipAdam(dec, lr=1e-4, grad_clip=5.0)
        enc = synthetic.features(Bb).to(dev)
        caps, lens = synthetic.captions(Bb, V, max_len=MAXLEN)
        caps = caps.to(dev)

        def step():
            preds, cs, dl, alphas = dec(enc, caps, lens)
            loss = attention_caption_loss(preds, cs, dl, alphas)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss
        name, batch = "configs[3]: glove_att, embed 300, fp64 fine-tuned embedding table, batch %d, T=24, V=9490" % Bb, Bb
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"metric": "train-step captions/s", "value": batch / (ms / 1e3), "unit": "captions/s", "n_gpus": 1,
                      "steps": args.steps, "ms_per_step": ms, "loss": float(loss.item()), "data": "synthetic",
                      "dtype": dec.precision, "config": {"workload": name}}), flush=True)


def bind_to_gpu_numa_node(device_index):
    """Pin this rank's threads (and hence its first-touched pinned host buffers) to the NUMA node its GPU hangs off: with
    one process per GPU the eight ranks otherwise all stage their H2D traffic through whichever node the OS picked."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None


def main():
    args = parse()
    if args.workload != "train" and args.impl != "reference":
        side_workload(args)
        return
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    assert torch.cuda.is_available(), "bench.py needs a GPU (the decoder has no CPU path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("ICD_BENCH_KEEP_NCCL_DEBUG") is None:
            os.environ["NCCL_DEBUG"] = "WARN"       # the NCCL version banner goes to stdout: keep stdout = ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()

    from icd_b200 import ops, synthetic
    from icd_b200._lib import lib
    from icd_b200.losses import attention_caption_loss
    from icd_b200.parallel import DataParallelClipAdam
    from icd_b200.vocabulary import synthetic_vocab
    import icd_b200.models.attention as my_att

    precision = args.precision
    if precision == "auto":
        precision = "bf16" if lib().icd_has_tensor_core_gemm() else "fp32"
    B = args.batch
    p = my_att.AttentionDecoderParams()
    p.vocab = synthetic_vocab(V)
    torch.manual_seed(0)
    dec = my_att.AttentionDecoder(dev, p)
    dec.fine_tune_embeddings(False)                      # basic_att: --fine_tune_embedding defaults to False
    dec = dec.to(dev)
    dec.precision = precision
    dec.train()
    opt = DataParallelClipAdam(dec, lr=1e-4, grad_clip=5.0)
    torch.manual_seed(1234 + rank)

    # pinned host copies of this rank's shard (synthetic; each rank a different seed) + resident device copies
    enc_h = synthetic.features(B, seed=1234 + rank).pin_memory()
    caps_h, lens = synthetic.captions(B, V, max_len=MAXLEN, seed=1234 + rank)
    caps_h = caps_h.pin_memory()
    enc_d = enc_h.to(dev, non_blocking=True)
    caps_d = caps_h.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def train_step(enc, caps):
        preds, caps_sorted, dl, alphas = dec(enc, caps, lens)
        loss = attention_caption_loss(preds, caps_sorted, dl, alphas, alpha_c=1.0, bf16_grad_only=True)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    uuid = None
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        pass
    sampler = ClockSampler(local_rank, uuid) if (rank == 0 and os.environ.get("ICD_BENCH_SAMPLER", "on") != "off") else None
    if sampler:
        sampler.wait_ready()
    n_warm = max(args.warmup, 6)                        # >= 3 required; 6 lets the caching allocator reach its steady state
    ops.prof_enable(True)                               # the library creates its timing events lazily: do that during warm-up
    for _ in range(n_warm):
        train_step(enc_d, caps_d)
    torch.cuda.synchronize()
    ops.prof_collect()                                  # discard the warm-up samples (the events stay allocated)
    import gc
    gc.collect()
    gc.disable()                    # a cyclic-GC pause of the launching thread inside the timed region starves the GPU queue
    ops.prof_enable(PROF_STRIDE)    # every PROF_STRIDE-th attention-step launch is bracketed by events (sampled: an event
    launches0 = ops.launch_count()  # record between two kernels costs ~4 us and suppresses their programmatic overlap)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()                       # LAST thing before the clock starts: the ranks leave it together (a gc.collect() between
                                    # the barrier and the first step skews them by milliseconds, paid in the first all-reduce)
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        loss = train_step(enc_d, caps_d)
        step_ev[i].record()
    ev1.record()
    t_issue = time.perf_counter() - t_wall0                      # host time to ISSUE the K steps (no sync inside)
    barrier()
    t_wall1 = time.perf_counter()
    per_step_ms = [round((ev0 if i == 0 else step_ev[i - 1]).elapsed_time(step_ev[i]), 3) for i in range(args.steps)]
    launches = ops.launch_count() - launches0
    prof = ops.prof_collect()
    ops.prof_enable(False)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    last_loss = float(loss.item())

    # ---- end-to-end timing: host buffers -> H2D -> step -> D2H loss, copies inside the timed region,
    #      next batch prefetched on a side stream while the current step computes ----
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()

        def e2e_run(enc_host, n_steps):
            bufs = [(torch.empty(enc_host.shape, dtype=enc_host.dtype, device=dev), torch.empty_like(caps_d)) for _ in range(2)]
            nocopy = bool(os.environ.get("ICD_BENCH_E2E_NOCOPY"))      # diagnostic only: isolates the cost of the H2D stream
            if nocopy:
                for b_ in bufs:                                        # valid contents once; the timed loop then copies nothing
                    b_[0].copy_(enc_host); b_[1].copy_(caps_h)
                torch.cuda.synchronize()
            ready = [torch.cuda.Event() for _ in range(2)]
            done = [torch.cuda.Event() for _ in range(2)]

            h2d_ev = []
            trace = []

            def prefetch(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done[i % 2])            # buffer free again?
                    e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e_a.record(copy_stream)
                    if not nocopy:
                        bufs[i % 2][0].copy_(enc_host, non_blocking=True)
                        bufs[i % 2][1].copy_(caps_h, non_blocking=True)
                    e_b.record(copy_stream)
                    h2d_ev.append((e_a, e_b))
                    ready[i % 2].record(copy_stream)

            loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()       # pinned landing buffer for the step losses
            loss_ready = [torch.cuda.Event() for _ in range(2)]

            def loop(n):
                """Every step: H2D of its inputs (prefetched one step ahead on the side stream) and a D2H read of its
                loss.  The loss is copied to pinned memory asynchronously and READ one step later, after the next step
                has been issued — the launch queue never drains (the reference's loss.item() right after backward
                stalls the pipeline every step)."""
                for d_ in done:
                    d_.record()
                prefetch(0)
                out = []
                for i in range(n):
                    if i + 1 < n:
                        prefetch(i + 1)
                    torch.cuda.current_stream().wait_event(ready[i % 2])
                    l = train_step(*bufs[i % 2])
                    done[i % 2].record()
                    ev = torch.cuda.Event(enable_timing=True); ev.record(); trace.append((time.perf_counter(), ev))
                    if i >= 1:                                   # read the PREVIOUS step's loss (its slot is reused at i+1)
                        loss_ready[(i - 1) % 2].synchronize()
                        out.append(float(loss_host[(i - 1) % 2]))
                    loss_host[i % 2:i % 2 + 1].copy_(l.detach().reshape(1), non_blocking=True)
                    loss_ready[i % 2].record()
                loss_ready[(n - 1) % 2].synchronize()
                out.append(float(loss_host[(n - 1) % 2]))
                assert len(out) == n
                return out[-1]
            loop(2)
            barrier()
            ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev2.record()
            loop(n_steps)
            ev3.record()
            barrier()
            ms2 = torch.tensor([ev2.elapsed_time(ev3)], device=dev)
            if world > 1:
                dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
            del bufs
            tr = trace[-n_steps:]
            deltas = sorted(tr[j][1].elapsed_time(tr[j + 1][1]) for j in range(len(tr) - 1))
            e2e_run.last_steady_ms = deltas[len(deltas) // 2] if deltas else None      # median step-to-step time
            e2e_run.last_first_ms = ev2.elapsed_time(tr[0][1]) if tr else None         # first step: its H2D copy is exposed
            if os.environ.get("ICD_BENCH_E2E_TRACE"):
                print("e2e trace: gpu step deltas (ms)", [round(tr[j][1].elapsed_time(tr[j + 1][1]), 2) for j in range(len(tr) - 1)],
                      "host issue deltas (ms)", [round(1e3 * (tr[j + 1][0] - tr[j][0]), 2) for j in range(len(tr) - 1)],
                      "ev2->first", round(ev2.elapsed_time(tr[0][1]), 2), "last->ev3", round(tr[-1][1].elapsed_time(ev3), 2), file=sys.stderr)
            h2d_ms = [a.elapsed_time(b) for a, b in h2d_ev[-n_steps:]]
            e2e_run.last_h2d_ms = sum(h2d_ms) / max(len(h2d_ms), 1)
            return float(ms2.item())

        # headline: features stored bf16 on the host (BASELINE.json configs[2] allows bf16-stored features; the bf16 tier
        # consumes them in place).  The fp32-host variant (exactly the reference's feature dtype) is reported beside it.
        host16 = precision == "bf16"
        enc_h16 = enc_h.to(torch.bfloat16).pin_memory() if host16 else None
        ms_main = e2e_run(enc_h16 if host16 else enc_h, args.steps)
        main_bytes = (enc_h16.numel() * 2 if host16 else enc_h.numel() * 4) + caps_h.numel() * 8
        e2e = {"value": B * world * args.steps / (ms_main / 1e3), "unit": "captions/s",
               "h2d_bytes_per_step": int(main_bytes), "d2h_bytes_per_step": 4,
               "ms_per_step": ms_main / args.steps,
               "host_feature_dtype": "bf16" if host16 else "fp32", "numa_node_rank0": numa_node,
               "h2d_ms_per_step": e2e_run.last_h2d_ms,
               "steady_ms_per_step": e2e_run.last_steady_ms, "first_step_ms": e2e_run.last_first_ms,
               "note": "per GPU: pinned host features + int64 captions copied H2D every step on a side stream "
                       "(double-buffered); every step's loss copied D2H to pinned memory and read on the host one step "
                       "later (pipelined read-back); the first step's copy cannot overlap anything and is inside the timed region "
                       "(first_step_ms), later copies hide behind the previous step (steady_ms_per_step)"}
        if host16:
            ms_f32 = e2e_run(enc_h, args.steps)
            e2e["fp32_host_features"] = {"value": B * world * args.steps / (ms_f32 / 1e3), "ms_per_step": ms_f32 / args.steps,
                                         "h2d_bytes_per_step": int(enc_h.numel() * 4 + caps_h.numel() * 8)}

    gc.enable()
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None
    if sampler:
        sampler.stop()

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        ATT_FWD_BYTES_PER_ROW, ATT_BWD_BYTES_PER_ROW = att_bytes_per_row(precision)
        fwd_s = prof["fwd_ms"] / 1e3
        achieved = (ATT_FWD_BYTES_PER_ROW * prof["fwd_rows"] / fwd_s / 1e9) if fwd_s > 0 else None
        bwd_s = prof["bwd_ms"] / 1e3
        achieved_bwd = (ATT_BWD_BYTES_PER_ROW * prof["bwd_rows"] / bwd_s / 1e9) if bwd_s > 0 else None
        line = {
            "metric": "attention-decoder train-step captions/s",
            "value": B * world * args.steps / (ms_total / 1e3), "unit": "captions/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(world, precision, B),
            "loss": last_loss,
            "per_step_ms": per_step_ms, "host_issue_ms_per_step": 1e3 * t_issue / args.steps,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "kernel": ("att_step_fwd_bf16_kernel" if precision == "bf16" else "att_step_fwd_kernel") +
                          " (fused additive-attention step, forward)",
                "feature_storage": "bf16" if precision == "bf16" else "fp32",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(),
                "peak_source": peak_src,
                "algorithmic_bytes_per_row": ATT_FWD_BYTES_PER_ROW,
                "launches": prof["fwd_launches"], "avg_launch_ms": (prof["fwd_ms"] / prof["fwd_launches"]) if prof["fwd_launches"] else None,
                "sampling": "every %d-th launch of the timed region is event-timed; share_of_step scales the sample back up" % PROF_STRIDE,
                "share_of_step": (PROF_STRIDE * prof["fwd_ms"] / ms_total) if ms_total else None,
                "bwd": {"kernel": "att_step_bwd_bf16_kernel" if precision == "bf16" else "att_step_bwd_kernel", "achieved": achieved_bwd,
                        "frac": (achieved_bwd / peak) if achieved_bwd else None,
                        "algorithmic_bytes_per_row": ATT_BWD_BYTES_PER_ROW,
                        "share_of_step": (PROF_STRIDE * prof["bwd_ms"] / ms_total) if ms_total else None},
            },
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_beam:
            try:
                line["beam5"] = beam_measure(dev, args.beam_images)
                line["beam5"]["note"] = "configs[4] on this GPU alone (images shard over GPUs with no collective)"
            except Exception as ex:          # never lose the headline line over the side measurement
                line["beam5"] = {"error": repr(ex)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_steps(steps=2, warmup=1)
            line["cpu_baseline"] = {"value": r["value"], "unit": "captions/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
