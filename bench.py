#!/usr/bin/env python
"""bench.py -- mel frames/s of the FCL-taco2 inference hot path on B200 (see DESIGN.md "Measurement").

  python bench.py --gpus 1 --steps 10 --warmup 3                       # our arm
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                                 # the reference's CPU path (oracle port)

A "step" = one pass of the whole hot path (encoder -> predictors -> length regulator -> decoder
-> postnet) over one synthetic LJSpeech-shaped batch. `value` = frames of all ranks / device time
(inputs resident in HBM); `e2e` = the same through model.inference_stream() with host buffers
(host planning, H2D of the inputs and D2H of the mels inside the timed region; the D2H of a batch overlaps
the compute of the next one, as in the decode driver).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout. Libraries chat on fd 1 (NCCL prints its version banner there), so fd 1 is
# pointed at stderr for the whole run and the JSON line is written to the saved, real stdout at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

from fcl_taco2_b200 import synth, hparams            # noqa: E402

# algorithmic work per unit (SURVEY.md 8(d) / BASELINE.md 3), MAC counts of the reference formulation
MACS = {
    "S": dict(decoder_row_step=1_438_720, postnet_frame=348_160, front_phoneme=3_593_856),
    "T": dict(decoder_row_step=15_941_632, postnet_frame=4_341_760, front_phoneme=8_611_968),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    In-process NVML (nvidia_ml_py) on a thread: a few light queries every 10 ms. An `nvidia-smi -lms` subprocess
    was measured to stall kernel launches for tens of ms when a poll (or its start-up) lands inside a timed step;
    it is only the fallback when NVML cannot be imported."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.samples = index, None, [], []
        self.nvml, self.handle, self.thread, self.run_flag, self.first = None, None, None, False, 0

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.run_flag = True
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _loop(self):
        n = self.nvml
        while self.run_flag:
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                try:
                    mem = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_MEM))
                except Exception:
                    mem = 0.0
                self.samples.append((mhz, reasons, mem))
            except Exception:
                pass
            time.sleep(0.01)

    def ready(self):
        return bool(self.samples) or bool(self.lines) or (self.nvml is None and self.proc is None)

    def mark(self):
        """Start of the timed region (the sampler is started before the warm-up)."""
        self.first = len(self.samples) if self.nvml is not None else len(self.lines)

    def stop(self):
        if self.nvml is not None:
            self.run_flag = False
            self.thread.join(timeout=1)
            n = self.nvml
            sel = self.samples[max(self.first - 1, 0):]
            names = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r, _m in sel))
            mhz = [m for m, _, _m in sel if m > 0]
            mem = [mm for _, _, mm in sel if mm > 0]
            return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "mem_mhz": float(np.median(mem)) if mem else None, "samples": len(sel), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for l in self.lines[max(self.first - 1, 0):]:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def hbm_copy_probe(dev, nbytes=1 << 30):
    """Device-to-device copy bandwidth of THIS box right before the run (read + write bytes, best of 5, CUDA events;
    outside every timed region): the HBM-bound stages (postnet, L2 flush, the decoder's scratch traffic) scale with it,
    so a slow step can be told from a slow box."""
    a = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    a.fill_(1)
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.copy_(a); e1.record()
        torch.cuda.synchronize(dev)
        best = max(best, 2.0 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del a, b
    return {"copy_gbs": best, "bytes": nbytes, "how": "torch b.copy_(a), read+write bytes, best of 6, before the timed region"}


def workload(args, rank):
    """LJSpeech-shaped synthetic batch for this rank; different seed per rank (weak scaling)."""
    xs, ds = synth.synth_batch(args.batch, seed=args.seed + 1000 * rank, stress=args.stress,
                               fixed_len=500 if args.stress else None)
    return xs, ds


def cpu_port_time(kind, xs, ds, budget_s, seed):
    """Time the oracle port (the reference's algorithm on torch CPU, per-utterance loop like
    tts.py:655-674) on a bounded sample. -> (frames/s, n_utts, frames, seconds, threads)"""
    from oracle import restate          # CPU baseline leg: bench.py touches oracle/ only here and in parity_check
    torch.set_num_threads(os.cpu_count())
    sd = synth.random_state_dict(hparams.preset(kind), seed)
    drop = restate.Dropout(0.5, 1, native=True)       # torch's F.dropout, as the reference runs it
    with torch.no_grad():
        restate.inference(sd, torch.from_numpy(xs[0]), dur=ds[0], dropout=drop, fast_lstm=True)     # warm-up
        t0 = time.perf_counter()
        frames = n = 0
        for x, d in zip(xs, ds):
            out = restate.inference(sd, torch.from_numpy(x), dur=d, dropout=drop, utt_index=n, fast_lstm=True)
            frames += out.shape[0]
            n += 1
            if time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    return frames / dt, n, frames, dt, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    xs, ds = workload(args, 0)
    vals, last = [], None
    for it in range(args.warmup + args.steps):
        last = cpu_port_time(args.model, xs, ds, budget_s=max(2.0, 60.0 / max(1, args.steps + args.warmup)), seed=args.seed)
        if it >= args.warmup:
            vals.append(last)
    fps = sum(v[2] for v in vals) / sum(v[3] for v in vals)
    sample = f"first {vals[-1][1]} utterances ({vals[-1][2]} frames) of the batch-{args.batch} workload per step, per-utterance loop"
    cfg = config_dict(args, world)
    cfg["precision"] = "fp32 (torch CPU; the GPU arm of the same workload computes its GEMMs with fp16 operands, fp32 accumulate)"
    line = {
        "metric": "mel frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(v[3] for v in vals) / len(vals), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": cfg,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": vals[-1][4], "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is Python and cannot travel to the GPU box: this is oracle/restate.py, the op-for-op "
                "torch-CPU restatement validated against the reference in the build container",
    }
    emit(line)


def config_dict(args, world):
    return {"workload": f"FCL-taco2-{args.model} batched inference, batch {args.batch} synthetic LJSpeech-shaped "
                        f"utterances per GPU" + (" (500-phoneme stress)" if args.stress else ""),
            "model_size": args.model, "per_gpu_batch": args.batch, "global_batch": args.batch * world,
            "precision": args.precision, "prenet_dropout": args.dropout, "forced_durations": True,
            "parallelism": f"utterance-sharded x{world}, no hot-path collective, final mel gather to rank 0"
                           + (" pipelined across steps (the mels of step i travel while step i+1 computes)" if world > 1 else ""),
            "l2": "flushed before every timed step (160 MiB write, INSIDE the timed region)"}


FLUSH_BYTES = 160 << 20          # > the 126 MB L2


class Arm:
    """One benchmark configuration on this rank: model + engine + uploaded batch."""

    def __init__(self, m, xs, ds, utt_ids=None, predicted=False):
        from fcl_taco2_b200 import plan as planmod
        self.m, self.eng, self.xs, self.ds = m, m.engine(), xs, ds
        self.utt_ids = utt_ids
        self.plan = planmod.make_plan(xs, None if predicted else ds, utt_ids=utt_ids)
        self.n_rows = self.plan.n_rows
        self.n_frames = int(sum(int(d.sum()) for d in ds)) if not predicted else None
        self.dinp, self.h2d_bytes = self.eng.upload(self.plan)
        self.bounds = None

    def chunk_bounds(self, k):
        from fcl_taco2_b200 import dist as fdist
        per_utt = np.add.reduceat(self.plan.dur.astype(np.int64), self.plan.utt_off[:-1].astype(np.int64))
        return fdist.chunk_bounds(np.concatenate([[0], np.cumsum(per_utt)]), k)


def time_arm(arm, steps, warmup, dropout, flush, dist=None, dev=None, gather=None, k_chunks=0, sampler=None):
    """W untimed + K timed passes of `arm`. The timed region is ONE bracket around the K passes (barrier +
    synchronize on both sides, CUDA events on the launching stream): L2 flush + pass (+ the hand-over of the mels to the
    pipelined gather) per step, and the wait for the last transfer at the end. Per-stage events inside give the
    breakdown. -> dict(total_ms, stage_ms (per step), launches (per step), frames (per step), last result)"""
    eng, m = arm.eng, arm.m

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step(timed):
        flush.fill_(1)
        eng.stage_events = [] if timed else None
        if gather is not None:
            res = eng.run_uploaded(arm.plan, arm.dinp, m.hp.zoneout_rate, dropout, 1, out_chunks=k_chunks,
                                   chunk_cb=gather.begin())
        else:
            res = eng.run_uploaded(arm.plan, arm.dinp, m.hp.zoneout_rate, dropout, 1)
        se = eng.stage_events
        eng.stage_events = None
        return res, se

    import gc
    # the warm-up keeps the previous result alive while the next pass runs, exactly like the timed loop below: the
    # caching allocator then owns BOTH output blocks (185 MB each) before the bracket opens. Without this the second
    # timed step found no free block and sat in cudaMalloc for several ms in the middle of the postnet launches (seen
    # as a one-off 9 ms postnet stage in ~1 run out of 6).
    wres = None
    for _ in range(warmup):
        wres, _se = one_step(False)
    del wres
    if gather is not None:
        gather.drain()
    gc.collect()
    gc.disable()            # a host GC pause between launches would show up as GPU idle time inside the bracket
    try:
        barrier()
        if sampler is not None:
            sampler.mark()
        launches0 = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stage_evs, res = [], None
        for _ in range(steps):
            res, se = one_step(True)
            stage_evs.append(se)
        bufs = None
        if gather is not None:
            t0 = torch.cuda.Event(enable_timing=True); t0.record()
            bufs = gather.drain()          # only the LAST step's transfer is still in flight here
            t1 = torch.cuda.Event(enable_timing=True); t1.record()
            stage_evs.append([("gather_tail_total", t0, t1)])
        e1.record()
        barrier()
    finally:
        gc.enable()
    stage_ms, per_step = {}, {}
    for se in stage_evs:
        for name, a, b in se:
            dt = a.elapsed_time(b)
            stage_ms[name] = stage_ms.get(name, 0.0) + dt / steps
            per_step.setdefault(name, []).append(dt)
    if os.environ.get("FCL_BENCH_DEBUG"):          # per-step stage times: tells a slow kernel from a one-off stall
        for name, v in per_step.items():
            print(f"[bench debug] {name}: " + " ".join(f"{x:.3f}" for x in v), flush=True)   # fd 1 is stderr here
    n_frames = arm.n_frames if arm.n_frames is not None else int(res.out.shape[0])
    return dict(total_ms=float(e0.elapsed_time(e1)), stage_ms=stage_ms, launches=(eng.launches - launches0) / steps,
                frames=n_frames, rows=arm.n_rows, res=res, bufs=bufs)


def parity_check(arm, res, dropout, seed, n=8):
    """Outside every timed region: `n` sampled utterances of the benchmark batch's LAST output against the oracle
    (oracle/restate.py, fp32 CPU, same counter-based dropout mask). bench.py uses oracle/ only as this checker and as
    the CPU baseline."""
    from oracle import restate
    outs = res.per_utterance()
    sd = {k: v.detach().cpu() for k, v in arm.m.state_dict().items()}
    rs = np.random.RandomState(1234)
    order = np.argsort([-len(x) for x in arm.xs], kind="stable")
    pick = sorted(set([int(order[0]), int(order[-1])] + rs.choice(len(arm.xs), min(n, len(arm.xs)), replace=False).tolist()))[:n]
    mx, l1, cnt = 0.0, 0.0, 0
    for i in pick:
        uid = i if arm.utt_ids is None else int(arm.utt_ids[i])
        ref = restate.inference(sd, torch.from_numpy(arm.xs[i]), dur=arm.ds[i], dropout=restate.Dropout(dropout, seed),
                                utt_index=uid, fast_lstm=True)
        diff = (outs[i].cpu() - ref).abs()
        mx = max(mx, float(diff.max()))
        l1 += float(diff.sum()); cnt += diff.numel()
    return {"max_abs": mx, "mean_l1": l1 / max(cnt, 1), "n": len(pick), "against": "oracle/restate.py fp32 (CPU), same Philox "
            "dropout mask; last output of the benchmark batch, checked outside the timed region"}


def decoder_roofline(model, t, pk, traffic=None):
    macs = MACS[model]
    dec_ms = t["stage_ms"].get("decoder_loop", 0.0)
    dec_flops = 2.0 * macs["decoder_row_step"] * t["frames"]
    achieved = dec_flops / (dec_ms * 1e-3) / 1e12 if dec_ms > 0 else 0.0
    return {"kernel": "decoder step loop (fcl_decoder_*), rank 0", "bound": "tensor", "achieved": achieved,
            "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": achieved / pk["tf_burst"], "traffic": traffic,
            "peak_source": pk["src"] + " bf16_tflops (burst: one ~1-2 ms launch inside a 4 ms step at full clocks)",
            "frac_of_sustained": achieved / pk["tf_sust"], "ms_per_launch": dec_ms,
            "algorithmic_flops_per_launch": dec_flops,
            "note": "algorithmic FLOPs = 2 x MAC per useful row-step (reference formulation, nothing hoisted) x frames"}


def secondary_configs(args, dev, flush, pk):
    """The other BASELINE.json configurations, driver-visible (value + decoder roofline fraction each): config 2
    (T, batch 32), config 4 (500-phoneme stress batch) and the production path with PREDICTED durations (biased
    duration head; the pass contains the one data-dependent D2H sync). 1 GPU, fewer steps."""
    from fcl_taco2_b200 import model as M
    out = {}
    steps, warm = max(3, min(args.steps, 8)), 3

    def entry(kind, t):
        r = decoder_roofline(kind, t, pk)
        return {"value": t["frames"] * steps / (t["total_ms"] * 1e-3), "unit": "frames/s", "ms_per_step": t["total_ms"] / steps,
                "frames_per_step": t["frames"], "decoder_ms": r["ms_per_launch"], "decoder_tflops": r["achieved"],
                "decoder_frac": r["frac"], "stage_ms_per_step": t["stage_ms"], "steps": steps}

    # config 2: teacher, batch 32
    mT = M.from_preset("T", seed=args.seed, device=dev, precision=args.precision).set_prenet_dropout(rate=args.dropout, seed=1)
    xs, ds = synth.synth_batch(32, seed=args.seed + 5)
    armT = Arm(mT, xs, ds)
    t = time_arm(armT, steps, warm, args.dropout, flush, dev=dev)
    out["T_batch32"] = entry("T", t)
    out["T_batch32"]["parity"] = parity_check(armT, t["res"], args.dropout, 1, n=3)
    del mT, armT, t
    torch.cuda.empty_cache()
    # config 4: stress, S, 64 utterances of 500 phonemes, skewed durations up to 40
    mS = M.from_preset("S", seed=args.seed, device=dev, precision=args.precision).set_prenet_dropout(rate=args.dropout, seed=1)
    xs, ds = synth.synth_batch(64, seed=args.seed + 6, stress=True, fixed_len=500)
    armS = Arm(mS, xs, ds)
    t = time_arm(armS, steps, warm, args.dropout, flush, dev=dev)
    out["S_stress_500x64"] = entry("S", t)
    out["S_stress_500x64"]["parity"] = parity_check(armS, t["res"], args.dropout, 1, n=2)
    del armS, t
    # production path: predicted durations (duration head biased so that a random-init model emits ~7 frames/phoneme)
    sd = dict(mS.state_dict())
    sd["duration_predictor.linear.bias"] = torch.full_like(sd["duration_predictor.linear.bias"], float(np.log(8.0)))
    mP = M.from_preset("S", seed=None, device="cpu", precision=args.precision)
    mP.load_state_dict(sd)
    mP = mP.to(dev).set_prenet_dropout(rate=args.dropout, seed=1)
    xs, ds = synth.synth_batch(args.batch, seed=args.seed)
    armP = Arm(mP, xs, ds, predicted=True)
    t = time_arm(armP, steps, warm, args.dropout, flush, dev=dev)
    out["S_predicted_durations"] = entry("S", t)
    out["S_predicted_durations"]["note"] = ("durations from the duration predictor (head bias log 8): includes the predictor, "
                                            "the serialised length regulator and the D2H sync of the frame totals")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="S", choices=["S", "T"])
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16 = 16-bit tensor-core path (bf16 = legacy alias; the operand format is a build property of the library)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--stress", action="store_true")
    ap.add_argument("--latency-utts", type=int, default=200)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip latency, secondary configs and the strong-scaling arm")
    ap.add_argument("--pair", type=int, default=-1, help="cta_group::2 decoder: 1 on, 0 off, -1 engine default")
    ap.add_argument("--pair-kernel", default="", choices=["", "v1", "v2"], help="cta_group::2 decoder variant (default: engine's)")
    ap.add_argument("--inflight", type=int, default=0, help="super-tiles in flight per CTA pair (v2 kernel; 0 = engine default)")
    ap.add_argument("--e2e-planner", type=int, default=0, help="inference_stream: plan batches in a helper thread (1) or inline (0)")
    ap.add_argument("--dropout", type=float, default=0.5, help="prenet dropout rate (reference default 0.5)")
    ap.add_argument("--gather", default="auto", choices=["auto", "nccl", "peer"], help="N > 1: transport of the final mel gather")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from fcl_taco2_b200 import model as M, plan as planmod, dist as fdist
    prev_affinity = fdist.bind_to_gpu_numa_node(local)     # pinned staging buffers land on the GPU's own NUMA node
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/fcl_nccl_%h_%p.log")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    dgroup = dist if world > 1 else None

    m = M.from_preset(args.model, seed=args.seed, device=dev, precision=args.precision)
    m.set_prenet_dropout(rate=args.dropout, seed=1)
    eng = m.engine()
    if args.pair >= 0:
        eng.use_pair = bool(args.pair)
    if args.pair_kernel:
        eng.pair_kernel = args.pair_kernel
    if args.inflight:
        eng.pair_inflight = args.inflight
    xs, ds = workload(args, rank)
    arm = Arm(m, xs, ds)
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    pk = peaks()
    hbm_probe = hbm_copy_probe(dev) if rank == 0 else None

    # groups of utterances whose mels are handed to the gather while the next group's postnet runs
    K_CHUNKS = 1 if world <= 2 else 4
    make_gather = lambda a: fdist.make_gather(a.chunk_bounds(K_CHUNKS), m.odim, dev, kind=args.gather) if world > 1 else None

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.ready() and time.time() - t_wait < 3.0:
            time.sleep(0.02)                       # the sampler is up and polling before the timed region starts
    gather = make_gather(arm)
    t = time_arm(arm, args.steps, args.warmup, args.dropout, flush, dgroup, dev, gather, K_CHUNKS, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    total_ms, stage_ms, n_frames, n_rows = t["total_ms"], t["stage_ms"], t["frames"], t["rows"]
    parity = parity_check(arm, t["res"], args.dropout, 1) if rank == 0 else None
    gather_kind = gather.kind if gather is not None else None
    if gather is not None:
        gather.close()
    del gather
    t["res"] = t["bufs"] = None

    # ---- e2e through the public API: host ids/durations in, host mels out. model.inference_stream() is the call a
    # decode driver makes: every batch is planned on the host, uploaded, decoded, and its mels are copied to pinned host
    # memory on a copy stream while the next batch computes. The timed region covers K whole batches, from the first
    # upload to the arrival of the last batch's mels on the host (L2 flushed before every batch, inside the region).
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def batches(k):
        for _ in range(k):
            yield {"xs": xs, "durs": ds}
    flush_l2 = lambda: flush.fill_(1)
    for _ in m.inference_stream(batches(6), before_batch=flush_l2, planner_thread=bool(args.e2e_planner)):
        pass
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    got = 0
    for outs in m.inference_stream(batches(args.steps), before_batch=flush_l2, planner_thread=bool(args.e2e_planner)):      # the generator returns a batch only when it is on the host
        got += len(outs)
    e1.record()
    barrier()
    assert got == args.steps * len(xs)
    e2e_ms = float(e0.elapsed_time(e1))

    # ---- strong scaling (BASELINE config 3 as written: ONE batch of `--batch` utterances sharded over the N GPUs)
    strong = None
    if world > 1 and not args.no_extra:
        gxs, gds = workload(args, 0)                                  # the same global batch on every rank
        shards = planmod.shard_utterances([int(d.sum()) for d in gds], world)
        mine = shards[rank]
        sarm = Arm(m, [gxs[i] for i in mine], [gds[i] for i in mine], utt_ids=mine)
        sg = make_gather(sarm)
        st = time_arm(sarm, args.steps, args.warmup, args.dropout, flush, dgroup, dev, sg, K_CHUNKS)
        ts = torch.tensor([st["total_ms"]], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        shard_equal = None
        if rank == 0:
            # determinism check: the gathered shards must equal, bit for bit, the whole batch decoded on ONE GPU
            full = m.inference_batch(gxs, durs=gds)
            shard_equal = True
            for r in range(world):
                rp = planmod.make_plan([gxs[i] for i in shards[r]], [gds[i] for i in shards[r]], utt_ids=shards[r])
                buf = st["res"].out if r == 0 else st["bufs"][r]
                fo = np.concatenate([[0], np.cumsum([int(gds[shards[r][int(j)]].sum()) for j in rp.perm])])
                for k, j in enumerate(rp.perm):
                    shard_equal = shard_equal and torch.equal(buf[int(fo[k]):int(fo[k + 1])], full[shards[r][int(j)]])
            del full
        tot_frames = float(sum(int(d.sum()) for d in gds))
        strong = {"scaling": "strong", "global_batch": args.batch, "value": tot_frames * args.steps / (float(ts[0]) * 1e-3),
                  "unit": "frames/s", "ms_per_step": float(ts[0]) / args.steps, "frames_per_step": tot_frames,
                  "shard_equal": shard_equal, "stage_ms_per_step_rank0": st["stage_ms"],
                  "note": "same kernels, one batch of --batch utterances LPT-sharded by frame count; gather to rank 0 pipelined across steps"}
        sg.close()
        del sg, st, sarm

    # ---- reduce over ranks: max time, sum frames
    tt = torch.tensor([total_ms, e2e_ms, float(n_frames), float(n_rows), float(t["launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, e2e_ms = float(tmax[0]), float(tmax[1])
        all_frames, all_rows, all_launches = float(tsum[2]), float(tsum[3]), float(tsum[4])
    else:
        all_frames, all_rows, all_launches = float(n_frames), float(n_rows), float(t["launches"])

    if rank == 0:
        value = all_frames * args.steps / (total_ms * 1e-3)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "decoder_traffic.json")))
            traffic = tj.get(f"{args.model}_{args.batch}_{args.precision}", {}).get("bytes")
        except Exception:
            pass
        line = {
            "metric": "mel frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else ("bf16" if eng.operand_format == "bf16" else "f16"), "data": "synthetic",
            "config": config_dict(args, world),
            "frames_per_step": all_frames, "phoneme_rows_per_step": all_rows,
            "clocks": clocks,
            "hbm_probe": hbm_probe,
            "e2e": {"value": all_frames * args.steps / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(arm.h2d_bytes), "d2h_bytes_per_step": int(n_frames * m.odim * 4),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(round(all_launches * args.steps)),
            "parity": parity,
            "roofline": decoder_roofline(args.model, t, pk, traffic),
            "stage_ms_per_step": stage_ms,
        }
        if gather_kind:
            line["gather"] = gather_kind
        if strong is not None:
            line["strong"] = strong
        # the HBM-side view SURVEY 8(d) asks for next to the tensor roofline: algorithmic bytes / stage time, rank 0
        sec = []
        for name, key, nbytes in (("postnet (first input + last output only: 640 B/frame)", "postnet", 640.0 * n_frames),
                                  ("length regulator scan + frame map (8 B/phoneme + 20 B/frame)", None,
                                   8.0 * n_rows + 20.0 * n_frames)):
            ms = stage_ms.get(key, 0.0) if key else stage_ms.get("len_reg", 0.0) + stage_ms.get("frame_map", 0.0)
            if ms > 0:
                gbs = nbytes / (ms * 1e-3) / 1e9
                sec.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                            "frac": gbs / pk["hbm"], "ms": ms})
        line["roofline_secondary"] = sec
        if not args.no_extra:
            line["p50_utt_latency_ms"] = latency_p50(m, args, dev)
            if world == 1 and not args.stress and args.model == "S":
                try:
                    line["secondary"] = secondary_configs(args, dev, flush, pk)
                except Exception as e:           # never lose the headline line to a secondary configuration
                    line["secondary"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only (torchrun pins OMP threads to 1)
            if prev_affinity:
                os.sched_setaffinity(0, prev_affinity)   # the CPU baseline gets every host core, not just the GPU's NUMA node
            fps, n, fr, dt, thr = cpu_port_time(args.model, xs, ds, budget_s=15.0, seed=args.seed)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": thr, "kind": "port",
                                    "sample": f"first {n} utterances ({fr} frames) of the same workload, per-utterance "
                                              f"loop of oracle/restate.py (torch CPU fp32), {dt:.1f} s"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def latency_p50(m, args, dev):
    """p50 single-utterance latency (batch = 1), device-timed through the public API with host input."""
    xs, ds = synth.synth_batch(args.latency_utts, seed=args.seed + 77)
    lat = []
    for i, (x, d) in enumerate(zip(xs, ds)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        m.inference(torch.from_numpy(x), None, dur=d)
        e1.record()
        e1.synchronize()
        if i >= 5:
            lat.append(e0.elapsed_time(e1))
    return float(np.median(lat))


if __name__ == "__main__":
    main()
