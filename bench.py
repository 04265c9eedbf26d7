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
region (+ the box's copy-only H2D ceiling measured in the same run).  `roofline`: the DOMINANT kernel of the step, the
fused attention-step backward (HBM-bound), algorithmic bytes / live CUDA-event time; the forward kernel nested under
`roofline.fwd`.  `beam5`: configs[4], image-sharded over all ranks with the final gather.  At N = 1 the line also carries
`other_configs` (configs[1] baseline decoder, configs[3] glove_att, configs[2] in length regime B = ragged captions with the
shrinking batch_size_t of models/attention.py:261-265, the fp32x3 / fp32 tiers of configs[2], the unmodified
reference run by eager PyTorch on the same GPU, a parity check of the timed path against the fp64 oracle) and
`cpu_baseline` / `cpu_baseline_configs0`.  `--impl reference` (and `cpu_baseline`): the UNMODIFIED reference modules from
baseline/_ref (baseline/install_ref.py) on the host cores, on a bounded 32-caption slice of the same workload.
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
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="icd_b200", choices=["icd_b200", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "fp32x3", "bf16"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="captions per GPU (default: the benchmark config)")
    ap.add_argument("--workload", default="train", choices=["train", "beam", "baseline", "glove", "tier", "ragged"],
                    help="train = BASELINE.json configs[2] (the metric's configuration, default); the others time the "
                         "remaining configs and print their own JSON line")
    ap.add_argument("--beam-images", type=int, default=1024, help="images per GPU for the beam-search measurement")
    ap.add_argument("--no-beam", action="store_true", help="skip the beam-5 side measurement of the default run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other configs / context figures / parity check of the N=1 line")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation on the host cores
# ---------------------------------------------------------------------------------------------------
REF_INSTALL = os.path.join(ROOT, "baseline", "_ref")


def _reference_namespace():
    """The UNMODIFIED reference, imported from its verbatim install under baseline/_ref (baseline/install_ref.py; git-ignored,
    shipped to the GPU box with the snapshot).  None when the install is absent."""
    if not os.path.isfile(os.path.join(REF_INSTALL, "models", "attention.py")):
        return None
    from oracle import reference_import
    return reference_import.load_reference(REF_INSTALL), reference_import


def cpu_reference_steps(steps, warmup, sample_b=CPU_SAMPLE_B, want="auto"):
    """Time `steps` train steps of the reference on a `sample_b`-caption slice of the benchmark batch, all host cores.
    kind "reference": the reference's own modules (models/attention.py:72-284, train_utils.py:2-12) driven by the body of its
    train loop (models/attention.py:396-430) restated literally; kind "port": oracle/decoders.py (only when the install is
    missing)."""
    import torch
    from torch.nn.utils.rnn import pack_padded_sequence
    from icd_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    enc = synthetic.features(sample_b)
    caps, lens = synthetic.captions(sample_b, V, max_len=MAXLEN)
    ref = _reference_namespace() if want in ("auto", "reference") else None
    times = []
    if ref is not None:
        ns, reference_import = ref
        kind = "reference"
        p = ns.AttentionDecoderParams()
        p.vocab = reference_import.make_reference_vocab(ns, V)
        torch.manual_seed(0)
        dec = ns.AttentionDecoder(torch.device("cpu"), p)
        dec.fine_tune_embeddings(False)                           # train.py:41 default
        dec.train()
        opt = torch.optim.Adam(params=filter(lambda q: q.requires_grad, dec.parameters()), lr=1e-4)   # :352-355
        criterion = torch.nn.CrossEntropyLoss()                   # :371
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            scores, caps_sorted, decode_lengths, alphas = dec(enc, caps, lens)                         # :396
            targets = caps_sorted[:, 1:]                                                               # :401
            scores = pack_padded_sequence(scores, decode_lengths, batch_first=True).data               # :405-406
            targets = pack_padded_sequence(targets, decode_lengths, batch_first=True).data             # :407-408
            loss = criterion(scores, targets)                                                          # :411
            loss += 1.0 * ((1. - alphas.sum(dim=1)) ** 2).mean()                                       # :414
            opt.zero_grad()                                                                            # :417
            loss.backward()                                                                            # :420
            ns.clip_gradient(opt, 5.0)                                                                 # :423
            opt.step()                                                                                 # :428
            loss.item()                                                                                # :433
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        kind = "port"
        from icd_b200.vocabulary import synthetic_vocab
        import icd_b200.models.attention as my_att
        from oracle import decoders as O
        p = my_att.AttentionDecoderParams()
        p.vocab = synthetic_vocab(V)
        torch.manual_seed(0)
        dec = my_att.AttentionDecoder(torch.device("cpu"), p)       # used as a seeded weight container only
        dec.fine_tune_embeddings(False)
        w = {k: v.detach().clone().requires_grad_(k != "embedding.weight") for k, v in dec.state_dict().items()}
        params = [t for t in w.values() if t.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-4)
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
    return dict(value=sample_b * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores, kind=kind,
                sample="%d-caption slice of the %d-caption batch (the reference's per-step enc_att recompute makes B=512 "
                       "a >10 s CPU step), T=24, V=%d, %d timed steps, %s" %
                       (sample_b, B_PER_GPU, V, len(times),
                        "unmodified reference modules from baseline/_ref" if kind == "reference" else "oracle port"))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup)
    cfg = workload_config(args.gpus, "fp32", B_PER_GPU)
    cfg["reference_arm_sample"] = r["sample"]
    cfg["reference_arm_batch"] = CPU_SAMPLE_B
    line = {
        "impl": "reference", "metric": "attention-decoder train-step captions/s", "value": r["value"],
        "unit": "captions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": "captions/s", "cores": r["cores"], "kind": r["kind"],
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
    """ncu `dram__bytes_read.sum + dram__bytes_write.sum` per launch of the two attention-step kernels, from the committed
    `ncu --set full` capture (profiles/roofline_traffic.json).  It cannot be measured inside a timed run (ncu replays every
    kernel ~40 times), so the bench line carries the committed figure and says where it comes from."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return {}
    return {}


def roofline_block(precision, prof, ms_total, peak, peak_src):
    """`roofline` of the bench line: the DOMINANT kernel of the step — the fused attention-step BACKWARD (largest share of
    the step) — with the forward kernel nested under "fwd".  achieved = algorithmic bytes per row x rows /
    live CUDA-event time of the sampled launches."""
    fwd_b, bwd_b = att_bytes_per_row(precision)
    tr = ncu_traffic()
    sfx = "_bf16_kernel" if precision == "bf16" else "_kernel"

    def one(direction, nbytes):
        ms, n, rows = prof[direction + "_ms"], prof[direction + "_launches"], prof[direction + "_rows"]
        ach = (nbytes * rows / (ms / 1e3) / 1e9) if ms > 0 else None
        return {"kernel": "att_step_%s%s (fused additive-attention step, %s)" % (direction, sfx, "backward" if direction == "bwd" else "forward"),
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": tr.get("att_step_%s_kernel_dram_bytes_per_launch" % direction),
                "traffic_source": tr.get("source", "profiles/roofline_traffic.json (committed ncu --set full capture, not this run)"),
                "algorithmic_bytes_per_row": nbytes, "rows_per_launch": (rows / n) if n else None,
                "launches": n, "avg_launch_ms": (ms / n) if n else None,
                "share_of_step": (PROF_STRIDE * ms / ms_total) if ms_total else None}
    blk = one("bwd", bwd_b)
    blk.update({"feature_storage": "bf16" if precision == "bf16" else "fp32", "peak_source": peak_src,
                "sampling": "every %d-th launch of the timed region is event-timed; share_of_step scales the sample back up" % PROF_STRIDE,
                "fwd": one("fwd", fwd_b)})
    return blk


def beam_measure(dev, n_img, k=5, max_steps=24, reps=2, world=1, rank=0):
    """BASELINE.json configs[4]: beam search k = 5 over synthetic features, caption cap 25 tokens (loop body runs for step
    = 1 .. max_steps + 1), weights from the beam fixture recipe so that captions terminate; `n_img` images
    PER GPU, image-sharded over `world` ranks with no collective until the final gather (gen_captions.beam_search_sharded).
    Timed region (CUDA events, max over ranks): the whole batched decode on every rank + the gather of lengths / token ids /
    scores to rank 0 + the device->host read there.  Afterwards (untimed) rank 0 re-decodes another rank's shard alone and
    checks the gathered captions are identical to that single-GPU run."""
    import torch
    import torch.distributed as dist
    from icd_b200.gen_captions import beam_search_batched, beam_search_sharded
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

    def shard(r):
        g = torch.Generator().manual_seed(77 + r)
        return torch.randn(n_img, 14, 14, C, generator=g).abs_().to(dev)
    feats = shard(rank)
    times, lens, seqs = [], None, None
    with torch.no_grad():
        for i in range(reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = beam_search_sharded(dec, feats, k, V - 3, V - 2, max_steps=max_steps)
            if rank == 0:
                lens = res["len"].cpu()
                seqs = res["seq"].cpu()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if i > 0:
                times.append(float(t.item()))
        check = None
        if rank == 0 and world > 1:                     # the gathered shard of the LAST rank == that shard decoded on one GPU
            r = world - 1
            one = beam_search_batched(dec, shard(r), k, V - 3, V - 2, max_steps=max_steps, want_alphas=False)
            check = bool(torch.equal(one["len"].cpu(), lens[r * n_img:(r + 1) * n_img]) and
                         torch.equal(one["seq"].cpu(), seqs[r * n_img:(r + 1) * n_img]))
    if rank != 0:
        return None
    ms = min(times)
    done = int((lens > 0).sum())
    total = n_img * world
    return {"metric": "beam-5 captions/s", "value": total / (ms / 1e3), "unit": "captions/s", "images": total,
            "images_per_gpu": n_img, "n_gpus": world, "beam": k,
            "max_caption_tokens": max_steps + 1, "ms": ms, "completed": done, "data": "synthetic",
            "gathered_equals_single_gpu_run": check,
            "precision": "fp32x3 (3-term bf16 split on tcgen05, fp32-grade logits; captions identical to the reference on the goldens)",
            "mean_len": float(lens[lens > 0].float().mean()) if done else 0.0,
            "config": {"workload": "configs[4]: beam search k=5, cap 25 tokens, %d images sharded over %d GPU(s), final gather of "
                                   "lengths + token ids + scores to rank 0 inside the timed region" % (total, world)}}


def _time_steps(step, steps, warmup):
    import torch
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(loss.item())


def measure_baseline(dev, precision="bf16", steps=20, warmup=5, world=1):
    """BASELINE.json configs[1]: baseline LSTM decoder, batch 128 (per GPU), L = 25, V = 9490, E = H = 512:
    fwd + loss + bwd + (gradient all-reduce) + clamp + Adam."""
    import torch
    from icd_b200 import synthetic
    from icd_b200.losses import baseline_caption_loss
    from icd_b200.parallel import DataParallelClipAdam
    import icd_b200.models.baseline as my_base
    Bb, L = 128, MAXLEN
    p = my_base.BaselineDecoderParams()
    p.vocab_size = V
    torch.manual_seed(0)
    dec = my_base.BaselineDecoder(p).to(dev)
    dec.precision = precision
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
    ms, loss = _time_steps(step, steps, warmup)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"metric": "baseline-decoder train-step captions/s", "value": Bb * world / (ms / 1e3), "unit": "captions/s",
            "n_gpus": world, "ms_per_step": ms, "steps": steps, "loss": loss, "dtype": precision, "data": "synthetic",
            "config": {"workload": "configs[1]: baseline LSTM decoder, batch 128/GPU, L=25, V=9490, E=H=512"}}


def measure_attention(dev, precision, batch, steps, warmup, glove=False, ragged=False):
    """The attention train step on one GPU in a given tier: configs[3] (glove_att: E = 300, fp64 fine-tuned table) when
    glove=True, else configs[2] in another precision tier (fp32x3 = the tier that meets the 1e-3 bar, fp32 = FMA tier).
    ragged=True: SURVEY.md section 8(d)'s length regime B (caption lengths U{8..25}, sorted descending: batch_size_t shrinks from B
    to a few rows over the 24 steps) instead of regime A (all captions 25 tokens, the reference's own training regime)."""
    import torch
    from icd_b200 import synthetic
    from icd_b200.losses import attention_caption_loss
    from icd_b200.parallel import DataParallelClipAdam
    from icd_b200.vocabulary import synthetic_vocab
    import icd_b200.models.attention as my_att
    p = my_att.AttentionDecoderParams()
    p.vocab = synthetic_vocab(V)
    if glove:
        p.embed_size = 300
    torch.manual_seed(0)
    dec = my_att.AttentionDecoder(dev, p)
    if glove:
        dec.load_pretrained_embeddins(synthetic.glove_like_table(V, 300))
    dec.fine_tune_embeddings(bool(glove))
    dec = dec.to(dev)
    dec.precision = precision
    dec.train()
    opt = DataParallelClipAdam(dec, lr=1e-4, grad_clip=5.0)
    enc = synthetic.features(batch).to(dev)
    caps, lens = synthetic.captions(batch, V, max_len=MAXLEN, lengths=("ragged" if ragged else None))
    caps = caps.to(dev)

    def step():
        preds, cs, dl, alphas = dec(enc, caps, lens)
        loss = attention_caption_loss(preds, cs, dl, alphas, bf16_grad_only=(precision == "bf16"))
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss
    ms, loss = _time_steps(step, steps, warmup)
    name = ("configs[3]: glove_att, embed 300, fp64 fine-tuned embedding table, batch %d, T=24, V=9490" % batch) if glove else \
           ("configs[2] in the %s tier: basic_att train step, batch %d, T=24, V=9490" % (precision, batch))
    out = {"metric": "attention-decoder train-step captions/s", "value": batch / (ms / 1e3), "unit": "captions/s", "n_gpus": 1,
           "ms_per_step": ms, "steps": steps, "loss": loss, "dtype": precision, "data": "synthetic", "config": {"workload": name}}
    if ragged:
        tokens = sum(l - 1 for l in lens)
        out["config"]["workload"] = ("configs[2], length regime B (SURVEY.md 8d): caption lengths U{8..25} sorted descending, batch %d, "
                                     "V=9490, %s tier" % (batch, precision))
        out["decode_tokens_per_step"] = tokens
        out["decode_tokens_per_s"] = tokens / (ms / 1e3)
        out["fraction_of_regime_A_tokens"] = tokens / float(batch * (MAXLEN - 1))
    return out


def measure_eager_torch_reference(dev, batch=B_PER_GPU, steps=2, warmup=1):
    """Context figure ("the bar to beat is eager PyTorch on the same B200"): the UNMODIFIED reference modules
    (baseline/_ref; the oracle port when the install is missing) run by eager PyTorch / cuBLAS on this GPU, same batch, same
    train-step body as the reference loop (models/attention.py:396-430).  Not a product path; nothing of icd_b200 runs here."""
    import torch
    from torch.nn.utils.rnn import pack_padded_sequence
    from icd_b200 import synthetic
    torch.backends.cuda.matmul.allow_tf32 = False          # the reference's defaults (SURVEY.md 8c caution vi)
    enc = synthetic.features(batch).to(dev)
    caps, lens = synthetic.captions(batch, V, max_len=MAXLEN)
    caps = caps.to(dev)
    ref = _reference_namespace()
    if ref is not None:
        ns, reference_import = ref
        p = ns.AttentionDecoderParams()
        p.vocab = reference_import.make_reference_vocab(ns, V)
        torch.manual_seed(0)
        dec = ns.AttentionDecoder(dev, p).to(dev)
        dec.fine_tune_embeddings(False)
        dec.train()
        opt = torch.optim.Adam(params=filter(lambda q: q.requires_grad, dec.parameters()), lr=1e-4)
        criterion = torch.nn.CrossEntropyLoss().to(dev)

        def step():
            scores, caps_sorted, decode_lengths, alphas = dec(enc, caps, lens)
            targets = caps_sorted[:, 1:]
            scores = pack_padded_sequence(scores, decode_lengths, batch_first=True).data
            targets = pack_padded_sequence(targets, decode_lengths, batch_first=True).data
            loss = criterion(scores, targets)
            loss += 1.0 * ((1. - alphas.sum(dim=1)) ** 2).mean()
            opt.zero_grad()
            loss.backward()
            ns.clip_gradient(opt, 5.0)
            opt.step()
            return loss
        kind = "unmodified reference modules (baseline/_ref), eager PyTorch fp32 on this GPU"
    else:
        return {"unavailable": "baseline/_ref not installed"}
    ms, loss = _time_steps(step, steps, warmup)
    return {"metric": "attention-decoder train-step captions/s", "value": batch / (ms / 1e3), "unit": "captions/s",
            "ms_per_step": ms, "steps": steps, "loss": loss, "dtype": "f32", "kind": kind,
            "config": {"workload": "configs[2] shape, batch %d, run by the reference's own code on the GPU" % batch}}


def side_workload(args):
    """`--workload baseline|glove|beam|tier`: the other BASELINE.json configs as their own JSON line (1 GPU, or N ranks
    under torchrun for baseline / beam: batch- resp. image-sharded)."""
    import torch
    import torch.distributed as dist
    import __graft_entry__
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    prec = "bf16" if args.precision in ("auto", "bf16") else args.precision
    if args.workload == "beam":
        line = beam_measure(dev, args.beam_images, world=world, rank=rank)
    elif args.workload == "baseline":
        line = measure_baseline(dev, prec, args.steps, max(args.warmup, 3), world=world)
    elif args.workload == "glove":
        line = measure_attention(dev, prec, args.batch, args.steps, max(args.warmup, 3), glove=True)
    elif args.workload == "ragged":
        line = measure_attention(dev, prec, args.batch, args.steps, max(args.warmup, 3), ragged=True)
    else:
        line = measure_attention(dev, prec, args.batch, args.steps, max(args.warmup, 3))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            time.sleep(0.5)              # NCCL's closing INFO lines first: the JSON result stays the last line on stdout
    if rank == 0:
        print(json.dumps(line), flush=True)


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
        # NCCL_DEBUG is left exactly as the caller set it (the driver reads the NCCL INFO lines to count the ranks of the
        # communicator); the JSON result is the LAST line rank 0 prints, written with one flush after the final barrier
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
    n_warm = args.warmup                                # exactly as asked (the timing rules want >= 3; the default is 3)
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
    host16, enc_h16 = precision == "bf16", None
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

    # ---- host -> device ceiling of this box: the SAME pinned buffers copied with nothing else running, all ranks at once ----
    if e2e is not None:
        src = enc_h16 if host16 else enc_h
        dst = torch.empty(src.shape, dtype=src.dtype, device=dev)
        dst.copy_(src, non_blocking=True)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        h1.record()
        barrier()
        t = torch.tensor([h0.elapsed_time(h1) / 5], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d_ms = float(t.item())
        nbytes = src.numel() * src.element_size()
        ceil = B * world / (h2d_ms / 1e3)
        e2e["h2d_ceiling"] = {"ms_per_batch_copy_alone": h2d_ms, "gbs_per_gpu": nbytes / (h2d_ms / 1e3) / 1e9,
                              "gbs_aggregate": world * nbytes / (h2d_ms / 1e3) / 1e9, "captions_per_s": ceil,
                              "note": "copy-only rate of the same pinned feature buffers, all %d rank(s) concurrently, max over "
                                      "ranks: an upper bound for any end-to-end number on this box" % world}
        e2e["frac_of_h2d_ceiling"] = e2e["value"] / ceil
        e2e["frac_of_resident"] = e2e["value"] / (B * world * args.steps / (ms_total / 1e3))
        del dst

    gc.enable()
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None
    if sampler:
        sampler.stop()

    # ---- configs[4]: beam search, image-sharded over ALL ranks with the final gather (collective: every rank takes part) ----
    beam5 = None
    if not args.no_beam:
        try:
            beam5 = beam_measure(dev, args.beam_images, world=world, rank=rank)
        except Exception as ex:              # never lose the headline line over the side measurement
            beam5 = {"error": repr(ex)[:200]}

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
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
            "roofline": roofline_block(precision, prof, ms_total, peak, peak_src),
        }
        if e2e:
            line["e2e"] = e2e
        if beam5 is not None:
            line["beam5"] = beam5
        if world == 1 and not args.no_extras:
            # the other BASELINE.json configs and context figures, one GPU, short runs (each guarded: the headline survives)
            extras = {}

            def guarded(name, fn):
                try:
                    extras[name] = fn()
                except Exception as ex:
                    extras[name] = {"error": repr(ex)[:300]}
                torch.cuda.empty_cache()
            guarded("parity_check", lambda: parity_check(dec, enc_d, caps_d, lens, dev))
            guarded("configs[1]_baseline_bf16", lambda: measure_baseline(dev, "bf16", steps=20, warmup=5))
            guarded("configs[1]_baseline_fp32x3", lambda: measure_baseline(dev, "fp32x3", steps=10, warmup=3))
            guarded("configs[3]_glove_bf16", lambda: measure_attention(dev, "bf16", B, steps=10, warmup=3, glove=True))
            guarded("configs[2]_regime_B_ragged_bf16", lambda: measure_attention(dev, "bf16", B, steps=10, warmup=3, ragged=True))
            guarded("configs[2]_fp32x3_tier", lambda: measure_attention(dev, "fp32x3", B, steps=5, warmup=3))
            guarded("configs[2]_fp32_tier", lambda: measure_attention(dev, "fp32", B, steps=3, warmup=2))
            guarded("eager_pytorch_reference_on_this_gpu", lambda: measure_eager_torch_reference(dev, B))
            line["other_configs"] = extras
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_steps(steps=2, warmup=1)
            line["cpu_baseline"] = {"value": r["value"], "unit": "captions/s", "cores": r["cores"], "kind": r["kind"],
                                    "sample": r["sample"]}
            r0 = cpu_reference_steps(steps=3, warmup=1, sample_b=4)
            line["cpu_baseline_configs0"] = {"value": r0["value"], "unit": "captions/s", "cores": r0["cores"], "kind": r0["kind"],
                                             "ms_per_step": r0["ms_per_step"],
                                             "sample": "configs[0] itself: basic_att, batch 4, 25 tokens, fwd + loss + bwd + clip + Adam "
                                                       "on the host cores (README.md:7), 3 timed steps"}
    if world > 1:
        # tear NCCL down BEFORE printing: with NCCL_DEBUG=INFO (left as the caller set it) the communicator prints while it
        # closes, and the JSON result must be the last line on stdout
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            time.sleep(0.5)              # let the other ranks' closing lines drain
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(line), flush=True)


def parity_check(dec, enc_d, caps_d, lens, dev, rows=32):
    """Checker, OUTSIDE every timed region: the exact module / tier / loss path that was just timed (current weights, train
    mode, a fixed dropout keep-mask) on the resident B-row batch, rows 0..31 compared with the fp64 oracle evaluated on the
    same 32 rows (rows are independent).  Full per-tensor gradient parity at this configuration: tests/test_gpu_timed_config.py."""
    import torch
    from oracle import decoders as O
    B, T, Dd = enc_d.shape[0], MAXLEN - 1, dec.decoder_dim
    g = torch.Generator().manual_seed(4321)
    keep = (torch.rand(T, B, Dd, generator=g) >= 0.5).to(torch.uint8)
    dec._dropout_mask_override = keep
    try:
        with torch.no_grad():
            preds, _, dl, alphas = dec(enc_d, caps_d, lens)
    finally:
        dec._dropout_mask_override = None
    w64 = {k: v.detach().cpu().double() for k, v in dec.state_dict().items()}
    masks = [keep[t, :rows].double() for t in range(T)]
    p64, _, dl64, a64 = O.attention_decoder_forward(w64, enc_d[:rows].cpu().double(), caps_d[:rows].cpu(), lens[:rows],
                                                    dropout_p=0.5, dropout_masks=masks, hoist=True)
    pr, al = preds[:rows].cpu().double(), alphas[:rows].cpu().double()
    top2 = p64.topk(2, dim=2).values
    flips = pr.argmax(2) != p64.argmax(2)
    return {"rows": rows, "of_batch": B, "tier": dec.precision,
            "logits_rel_err": float((pr - p64).norm() / p64.norm()), "alphas_rel_err": float((al - a64).norm() / a64.norm()),
            "greedy_id_flips": int(flips.sum()), "positions": int(flips.numel()),
            "greedy_id_flips_where_margin_gt_2e-2": int((flips & ((top2[..., 0] - top2[..., 1]) > 2e-2)).sum()),
            "stated_tolerance": {"logits": 5e-3, "alphas": 5e-3} if dec.precision == "bf16" else {"logits": 1e-4, "alphas": 1e-4},
            "oracle": "oracle/decoders.py in fp64 on the same rows (checker only)"}


if __name__ == "__main__":
    main()
